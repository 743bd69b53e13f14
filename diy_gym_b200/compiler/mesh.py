"""Vertex readers for the three mesh formats the vendored URDFs reference (binary/ASCII STL, COLLADA,
Wavefront OBJ) and the reduction of a vertex cloud to a collision/visual proxy.

Only vertex positions are needed: collision meshes become convex hulls in the reference's physics
engine, and this backend replaces each hull by a fitted capsule or box (SURVEY.md §7.3 item 6 - a
documented deviation; contact parity is distributional, single-step parity is contact-free).
"""
import os
import re
import struct
import xml.etree.ElementTree as ET

import numpy as np


def _read_stl(path):
    with open(path, 'rb') as f:
        data = f.read()
    if len(data) >= 84:
        ntri = struct.unpack_from('<I', data, 80)[0]
        if 84 + 50 * ntri == len(data):
            rec = np.frombuffer(data, dtype=np.dtype([('n', '<f4', 3), ('v', '<f4', (3, 3)), ('a', '<u2')]), count=ntri,
                                offset=84)
            return rec['v'].reshape(-1, 3).astype(np.float64)
    text = data.decode('ascii', errors='ignore')
    verts = re.findall(r'vertex\s+([-\d.eE+]+)\s+([-\d.eE+]+)\s+([-\d.eE+]+)', text)
    return np.array(verts, dtype=np.float64)


def _read_obj(path):
    verts = []
    with open(path, 'r', errors='ignore') as f:
        for line in f:
            if line.startswith('v '):
                verts.append([float(t) for t in line.split()[1:4]])
    return np.array(verts, dtype=np.float64)


def _read_dae(path):
    root = ET.parse(path).getroot()
    ns = ''
    if root.tag.startswith('{'):
        ns = root.tag[:root.tag.index('}') + 1]
    out = []
    for mesh in root.iter(ns + 'mesh'):
        pos_id = None
        for vert in mesh.iter(ns + 'vertices'):
            for inp in vert.iter(ns + 'input'):
                if inp.get('semantic') == 'POSITION':
                    pos_id = inp.get('source', '').lstrip('#')
        for src in mesh.iter(ns + 'source'):
            if pos_id is not None and src.get('id') != pos_id:
                continue
            if pos_id is None and 'position' not in (src.get('id') or '').lower():
                continue
            fa = src.find(ns + 'float_array')
            if fa is not None and fa.text:
                out.append(np.array(fa.text.split(), dtype=np.float64).reshape(-1, 3))
    if not out:
        return np.zeros((0, 3))
    return np.concatenate(out, axis=0)


def load_vertices(path):
    """Return the (V,3) vertex cloud of a mesh file (no up-axis conversion: the reference's importer
    ignores the COLLADA up-axis tag for collision geometry)."""
    ext = os.path.splitext(path)[1].lower()
    if ext == '.stl':
        return _read_stl(path)
    if ext == '.obj':
        return _read_obj(path)
    if ext == '.dae':
        return _read_dae(path)
    raise ValueError('Unsupported mesh format: ' + path)


def fit_proxy(verts):
    """Fit a collision proxy to a vertex cloud given in the geometry frame.

    Returns dict(type, dims, center, axis) where `axis` is the index of the capsule axis (capsules are
    canonically along local z; the caller rotates).  An exact 8-corner box stays a box; elongated
    roughly-round clouds become capsules; everything else becomes its axis-aligned bounding box.
    """
    lo, hi = verts.min(axis=0), verts.max(axis=0)
    center = 0.5 * (lo + hi)
    ext = hi - lo
    uniq = np.unique(np.round(verts, 9), axis=0)
    if len(uniq) == 8:
        corners = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
        d = np.abs(uniq[:, None, :] - corners[None, :, :]).sum(-1).min(axis=1)
        if np.all(d < 1e-6 * max(1.0, ext.max())):
            return dict(type='box', dims=(0.5 * ext).tolist(), center=center.tolist(), axis=2, half=(0.5 * ext).tolist(), exact=True)
    order = np.argsort(ext)
    a, b, c = ext[order]
    if c > 0 and b > 0 and (b - a) <= 0.25 * b and c >= 1.05 * b:
        radius = 0.25 * (a + b)
        half = max(0.5 * c - radius, 0.0)
        return dict(type='capsule', dims=[float(radius), float(half), 0.0], center=center.tolist(), axis=int(order[2]),
                    half=(0.5 * ext).tolist())
    return dict(type='box', dims=(0.5 * ext).tolist(), center=center.tolist(), axis=2, half=(0.5 * ext).tolist())


def reduced_hull(verts, max_verts=32, max_planes=64):
    """Convex hull of a vertex cloud reduced to at most `max_verts` vertices (SURVEY hard part 6): the support points of the cloud
    along a fixed fan of directions (the six axes + a Fibonacci sphere), thinned by farthest-point selection - an INNER
    approximation of the true hull (pybullet collides the full hull, 72-1500 vertices for the vendored meshes).
    Returns (V [n, 3], P [m, 4]) with the half-spaces  P[:, :3] . x <= P[:, 3]  of the reduced hull, or None if the cloud is flat.
    Needs scipy (build container / asset compilation only: the result is stored in the compiled model descriptor)."""
    from scipy.spatial import ConvexHull, QhullError
    verts = np.asarray(verts, np.float64)
    if len(verts) < 4:
        return None
    try:
        full = ConvexHull(verts)
    except QhullError:
        return None
    hv = verts[full.vertices]
    if len(hv) > max_verts:
        n = 96
        k = np.arange(n) + 0.5
        phi, th = np.arccos(1 - 2 * k / n), np.pi * (1 + 5 ** 0.5) * k
        dirs = np.concatenate([np.eye(3), -np.eye(3), np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1)])
        idx = np.unique(np.argmax(hv @ dirs.T, axis=0))
        cand = hv[idx]
        if len(cand) > max_verts:   # farthest-point thinning, seeded with the extreme points along the axes
            keep = list(np.unique(np.concatenate([np.argmax(cand, axis=0), np.argmin(cand, axis=0)])))
            d = np.min(np.linalg.norm(cand[:, None, :] - cand[keep][None, :, :], axis=2), axis=1)
            while len(keep) < max_verts:
                j = int(np.argmax(d))
                keep.append(j)
                d = np.minimum(d, np.linalg.norm(cand - cand[j], axis=1))
            cand = cand[sorted(keep)]
        hv = cand
    try:
        red = ConvexHull(hv)
    except QhullError:
        return None
    V = hv[red.vertices]
    eq = red.equations                                   # n . x + d <= 0 inside
    planes = np.concatenate([eq[:, :3], -eq[:, 3:4]], axis=1)
    key = np.round(planes / np.maximum(np.abs(planes).max(), 1e-12), 5)
    _, first = np.unique(key, axis=0, return_index=True)
    planes = planes[np.sort(first)]
    if len(planes) > max_planes:                         # keep the largest facets' planes: merge by dropping near-duplicates of kept normals
        order = np.argsort(-np.array([np.sum(np.isclose(eq[:, :3] @ pl[:3], 1.0, atol=1e-3)) for pl in planes]))
        planes = planes[np.sort(order[:max_planes])]
    return V, planes

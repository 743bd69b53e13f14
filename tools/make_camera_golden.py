#!/usr/bin/env python
"""Writes tests/golden/refcamera_<config>.npz: what the REFERENCE's own camera add-on returns after reset.

The reference's `DIYGym` and `Camera` classes (/root/reference/diy_gym, unmodified, imported in the build container only) run on
the oracle-backed pybullet shim, whose getCameraImage renders from NOTHING BUT the view and projection matrices the add-on hands
it (oracle/bullet_restatement.c: dgo_get_camera_image) and returns what pybullet returns (rgba bytes, non-linear depth buffer).
So the recorded `rgb` / `depth` carry the reference's conventions end to end: T_world_cam = inv(T_world_parent T_parent_cam) from
getLinkState[4:6] (camera.py:58-68), fov / aspect / clipping (camera.py:29-45), row order and (W, H) labelling (camera.py:76-82),
rgb = bytes / 255, eye-space depth from the depth buffer (camera.py:80-85).  tests/test_camera_reference.py renders the same
scenes with the compiled camera (oracle ray caster on CPU, CUDA kernel on the GPU) and compares.
The pixels themselves come from this repo's ray caster - TinyRenderer is absent - so this pins conventions, not shading."""
import os
import sys
import tempfile

import numpy as np
import yaml

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shim'))
sys.path.insert(0, REF)

from diy_gym import DIYGym  # noqa: E402  (the reference)

CONFIGS = {'from_the_readme': os.path.join(REF, 'examples', 'from_the_readme', 'from_the_readme.yaml'),
           'basic_env': os.path.join(REF, 'diy_gym', 'tests', 'basic_env.yaml')}


def cameras(obs, prefix=''):
    for k, v in obs.items():
        if isinstance(v, dict):
            if 'rgb' in v:
                yield prefix + '/' + k, v
            else:
                yield from cameras(v, prefix + '/' + k)


def main():
    for name, path in CONFIGS.items():
        node = yaml.load(open(path), Loader=yaml.FullLoader)
        node['render'] = False
        tmp = os.path.join(tempfile.mkdtemp(), name + '.yaml')
        yaml.dump(node, open(tmp, 'w'))
        env = DIYGym(tmp)
        obs = env.reset()
        rec = {}
        for key, cam in cameras(obs):
            rgb = np.asarray(cam['rgb'])
            rec['key'] = np.array([key])
            rec['rgb_u8'] = np.round(rgb * 255).astype(np.uint8)          # the reference's values are bytes / 255 exactly
            rec['depth'] = np.asarray(cam['depth'], np.float32)
            rec['shape_rgb'], rec['shape_depth'] = np.array(rgb.shape), np.array(np.asarray(cam['depth']).shape)
        dst = os.path.join(ROOT, 'tests', 'golden', 'refcamera_' + name + '.npz')
        np.savez_compressed(dst, **rec)
        print('%-18s %s rgb %s depth %s [%.3f, %.3f] -> %.1f KB' % (name, rec['key'][0], tuple(rec['shape_rgb']), tuple(rec['shape_depth']),
                                                                    rec['depth'].min(), rec['depth'].max(), os.path.getsize(dst) / 1024))
        env.close()


if __name__ == '__main__':
    main()

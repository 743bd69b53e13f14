"""Aggregate `ncu --page source --print-source cuda,sass --csv` by CUDA source line and by enclosing function.
Usage: python tools/ncu_cuda_lines.py <prof.ncu-rep> [top]   (the report must be captured with --import-source on)"""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
files = {}     # file -> {line: [samples, inst, thread inst, source]}
cur, hdr = None, None
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = files.setdefault(r[1], {})
        continue
    if r and r[0] == 'Line No':
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if cur is None or hdr is None or not r or not r[0].strip().isdigit():
        continue
    try:
        s, i, t = int(r[hdr['# Samples']] or 0), int(r[hdr['Instructions Executed']] or 0), int(r[hdr['Thread Instructions Executed']] or 0)
    except (ValueError, IndexError):
        continue
    e = cur.setdefault(int(r[0]), [0, 0, 0, r[1]])
    e[0] += s; e[1] += i; e[2] += t
ts = sum(e[0] for f in files.values() for e in f.values()) or 1
ti = sum(e[1] for f in files.values() for e in f.values()) or 1
print('total samples %d, warp instructions %d' % (ts, ti))
byfn = collections.Counter(); byfn_i = collections.Counter()
allrows = []
import os
def fn_table(path):
    """line -> enclosing function, from the source file on disk (falls back to '?')"""
    tab, fn = {}, '?'
    local = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'diy_gym_b200', 'csrc', os.path.basename(path))
    if not os.path.isfile(local):
        return tab
    for i, l in enumerate(open(local).read().splitlines(), 1):
        if re.match(r'^(DG_FN|DG_HD|DG_NOINLINE|__device__|__global__|template <[^>]*> (DG_FN|DG_HD|__device__|static))', l) and '(' in l:
            m = re.search(r'(\w+)\(', l)
            if m:
                fn = m.group(1)
        tab[i] = fn
    return tab
for path, lines in files.items():
    tab = fn_table(path)
    for n in sorted(lines):
        src = lines[n][3]
        fn = tab.get(n, os.path.basename(path))
        byfn[fn] += lines[n][0]; byfn_i[fn] += lines[n][1]
        allrows.append((lines[n][0], lines[n][1], lines[n][2], path.split('/')[-1], n, fn, src.strip()[:100]))
print('--- by function')
for f, s in byfn.most_common(25):
    print('%5.1f%% samp %5.1f%% inst  %s' % (100.0 * s / ts, 100.0 * byfn_i[f] / ti, f))
print('--- by line')
for s, i, t, p, n, fn, src in sorted(allrows, reverse=True)[:top]:
    print('%5.1f%% samp %5.1f%% inst thr/inst %4.1f  %s:%d [%s]  %s' % (100.0 * s / ts, 100.0 * i / ti, t / max(i, 1), p, n, fn, src))

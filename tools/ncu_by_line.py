"""Aggregate an ncu SASS source page (csv) by CUDA source line using nvdisasm line info.
Usage: python tools/ncu_by_line.py <prof.ncu-rep> <lib.so> <kernel substring> [top]"""
import csv
import collections
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
dis = subprocess.run(['nvdisasm', '--print-line-info', '-c', cubin], capture_output=True, text=True).stdout.splitlines()
# offsets -> (file:line of innermost, outermost inline chain)
line_of = {}
in_k = False
cur = None
for l in dis:
    if l.startswith('\t.section\t.text.'):
        in_k = kname in l
        cur = None
        continue
    if not in_k:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = rows[1]
ia, isamp, iinst, ithr = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed')
base = int(rows[2][ia], 16)
samp, inst, thr = collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[2:]:
    if len(r) <= ithr:
        continue
    off = int(r[ia], 16) - base
    key = line_of.get(off, ('?', 0))
    samp[key] += int(r[isamp]); inst[key] += int(r[iinst]); thr[key] += int(r[ithr])
ts, ti = sum(samp.values()), sum(inst.values())
print('total samples %d, warp instructions %d, thread instr %d' % (ts, ti, sum(thr.values())))
src_cache = {}
SRC_DIRS = [os.environ['NCU_SRC']] if os.environ.get('NCU_SRC') else []
SRC_DIRS += [os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', d) for d in ('diy_gym_b200/csrc', 'include')]


def src(f, n):
    for d in SRC_DIRS:
        p = os.path.join(d, f)
        if os.path.isfile(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][n - 1].strip()[:110] if 0 < n <= len(src_cache[p]) else ''
    return ''
for key, s in samp.most_common(top):
    print('%5.1f%% samp %5.1f%% inst  thr/inst %4.1f  %s:%d  %s' % (100.0 * s / ts, 100.0 * inst[key] / ti, thr[key] / max(inst[key], 1), key[0], key[1], src(*key)))

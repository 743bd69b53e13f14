"""Aggregate an ncu SASS source page by enclosing source function, with the stall reasons.
Usage: NCU_SRC=<dir of the sources the .so was built from> python tools/ncu_by_func.py <prof.ncu-rep> <lib.so> <kernel substring>"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kname = sys.argv[1], sys.argv[2], sys.argv[3]
srcdir = os.environ.get('NCU_SRC') or os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'diy_gym_b200', 'csrc')
tmp = tempfile.mkdtemp()
subprocess.check_call(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
dis = []
for f in sorted(os.listdir(tmp)):   # one cubin per compiled object (diy_gym_b200/build.py): take the one that holds the kernel
    if f.endswith('.cubin'):
        out_ = subprocess.run(['nvdisasm', '--print-line-info', '-c', os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if ('.text.' in out_) and any(kname in l for l in out_.splitlines() if l.startswith('\t.section\t.text.')):
            dis = out_.splitlines()
            break
# innermost line + the full inline chain (outermost function attribution)
line_of, in_k, cur = {}, False, None
for l in dis:
    if l.startswith('\t.section\t.text.'):
        in_k, cur = kname in l, None
        continue
    if not in_k:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if m.group(3):   # attribute helpers (dg_math.cuh) to the caller's line
            f, n = (m.group(3), int(m.group(4))) if 'dg_math' in m.group(1) or 'intrinsics' in m.group(1) else (m.group(1), int(m.group(2)))
        else:
            f, n = m.group(1), int(m.group(2))
        cur = (os.path.basename(f), n)
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
# function of each line of dg_env.cuh
fn = {}
for fname in ('dg_env.cuh', 'dg_kernels.cu'):
    p = os.path.join(srcdir, fname)
    if not os.path.isfile(p):
        continue
    cur = '?'
    for i, l in enumerate(open(p).read().splitlines(), 1):
        m = re.match(r'^(?:template <[^>]*>\s*)?(?:DG_NOINLINE )?(?:DG_FN|DG_HD|__global__|__device__ __forceinline__)[\w \*<>]*?\**(\w+)\(', l)
        if m:
            cur = m.group(1)
        fn[(fname, i)] = cur
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
base = int(rows[2][col['Address']], 16)
keys = ['# Samples', 'Instructions Executed', 'Thread Instructions Executed', 'stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_no_inst', 'stall_branch_resolving', 'stall_barrier', 'stall_math', 'stall_lg', 'stall_mio']
agg = collections.defaultdict(lambda: collections.Counter())
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    key = line_of.get(int(r[col['Address']], 16) - base, ('?', 0))
    f = fn.get(key, key[0])
    for k in keys:
        agg[f][k] += int(r[col[k]] or 0)
tot = collections.Counter()
for f in agg:
    tot.update(agg[f])
print('total: samples %d, warp instr %d, thread instr %d' % (tot['# Samples'], tot['Instructions Executed'], tot['Thread Instructions Executed']))
print('%-22s %6s %6s %5s | %s' % ('function', 'samp%', 'inst%', 'thr', '  '.join(k.replace('stall_', '')[:8].rjust(8) for k in keys[3:])))
for f, c in sorted(agg.items(), key=lambda kv: -kv[1]['# Samples'])[:22]:
    print('%-22s %6.1f %6.1f %5.1f | %s' % (f[:22], 100.0 * c['# Samples'] / tot['# Samples'], 100.0 * c['Instructions Executed'] / tot['Instructions Executed'],
                                          c['Thread Instructions Executed'] / max(c['Instructions Executed'], 1),
                                          '  '.join(('%.1f' % (100.0 * c[k] / tot['# Samples'])).rjust(8) for k in keys[3:])))

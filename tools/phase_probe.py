"""Per-phase cycle breakdown of one step-kernel launch (dg_debug_phase_cycles): where a block's time goes, for the average
block and for the slowest one.  Usage: python tools/phase_probe.py <config> [n_envs] [presteps] [steps measured]"""
import ctypes
import os
import re
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from bench import CONFIGS, action_ranges, register_example_addons  # noqa: E402
from diy_gym_b200 import DIYGym  # noqa: E402

register_example_addons()
name = sys.argv[1] if len(sys.argv) > 1 else 'r2d2_maze'
n = int(sys.argv[2]) if len(sys.argv) > 2 else CONFIGS[name][1]
pre = int(sys.argv[3]) if len(sys.argv) > 3 else 60
meas = int(sys.argv[4]) if len(sys.argv) > 4 else 4
env = DIYGym(os.path.join(ROOT, CONFIGS[name][0]), num_envs=n)
w = env.world
lo, hi = action_ranges(env)
lo, hi = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
for _ in range(pre):
    if w.n_act:
        w.action.copy_(lo + (hi - lo) * torch.rand((n, w.n_act), device='cuda'))
    w.step()
torch.cuda.synchronize()
grid = w.grid_blocks
assert w.L.dg_debug_phase_cycles(w._h, 1) == 0
for _ in range(meas):
    w.step()
buf = np.zeros(grid * 64, np.uint64)
assert w.L.dg_debug_read(w._h, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), buf.size) == 0
w.L.dg_debug_phase_cycles(w._h, 0)
t = buf.reshape(grid, 64).astype(np.float64) / meas
src = open(os.path.join(ROOT, 'diy_gym_b200', 'csrc', 'dg_env.cuh')).read().splitlines()
label = {}
for i, l in enumerate(src, 1):
    if 'DG_PHASE(' in l and '#define' not in l or 'rs_solve_block<' in l:
        m = re.search(r'DG_PHASE\((?:if \([^)]*\)[^;]*?)?\s*(\w+)\(', l)
        label.setdefault(i & 63, []).append((m.group(1) if m else l.strip()[:40]) + ':%d' % i)
for i, l in enumerate(src, 1):   # the solver's own timing line
    if 'C.dbg[(size_t)blockIdx.x * 64 + (__LINE__ & 63)] +=' in l and 'define' not in src[i - 2] and 'DG_PHASE' not in l:
        label.setdefault(i & 63, []).append('rs_solve_block:%d' % i)
tot = t.sum(axis=1)
slow = int(np.argmax(tot))
clk = 1.965e9
print('%s: %d envs, team %d, %d blocks x %d threads; per-launch block time mean %.3f ms, max %.3f ms (block %d), min %.3f ms' %
      (name, n, w.team, grid, w.block_threads, tot.mean() / clk * 1e3, tot.max() / clk * 1e3, slow, tot.min() / clk * 1e3))
print('%-58s %10s %10s %10s' % ('phase (source line)', 'mean ms', 'slowest ms', 'p95 ms'))
order = np.argsort(-t.mean(axis=0))
for k in order:
    if t[:, k].max() == 0:
        continue
    print('%-58s %10.4f %10.4f %10.4f' % (' | '.join(label.get(int(k), ['?'])), t[:, k].mean() / clk * 1e3, t[slow, k] / clk * 1e3, np.percentile(t[:, k], 95) / clk * 1e3))

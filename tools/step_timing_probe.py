"""Back-to-back step time (no L2 flush) against the bench's cold-L2 step time, and the workspace sizes behind the difference.
Usage: python tools/step_timing_probe.py <config> [n_envs] [presteps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from bench import CONFIGS, ROOT, action_ranges, register_example_addons  # noqa: E402
from diy_gym_b200 import DIYGym  # noqa: E402

register_example_addons()
name = sys.argv[1] if len(sys.argv) > 1 else 'r2d2_maze'
n = int(sys.argv[2]) if len(sys.argv) > 2 else CONFIGS[name][1]
pre = int(sys.argv[3]) if len(sys.argv) > 3 else 300
env = DIYGym(os.path.join(ROOT, CONFIGS[name][0]), num_envs=n)
w = env.world
lo, hi = action_ranges(env)
lo, hi = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
acts = [lo + (hi - lo) * torch.rand((n, w.n_act), device='cuda') for _ in range(8)] if w.n_act else None


def step(i):
    if acts is not None:
        w.action.copy_(acts[i % 8])
    w.step()


for i in range(pre):
    step(i)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
out = {}
for mode in ('back_to_back', 'l2_flushed'):
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(40)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(40)]
    for i in range(40):
        if mode == 'l2_flushed':
            flush.fill_(i & 255)
        e0[i].record()
        step(i)
        e1[i].record()
    torch.cuda.synchronize()
    out[mode] = sum(a.elapsed_time(b) for a, b in zip(e0, e1)) / 40
S, P = env.scene.hdr['S'] if 'S' in env.scene.hdr else w.state.shape[1], w.param.shape[1]
print('%s n %d split %s: %.3f ms per step back to back, %.3f ms with the L2 flushed before every step' % (name, n, w.split, out['back_to_back'], out['l2_flushed']))
print('  per environment: state row %d B, hot workspace (carry) %d B; state + carry of the batch: %.1f MB' %
      (w.state.shape[1] * 4, w.ws_floats * 4, n * (w.state.shape[1] + w.ws_floats) * 4 / 1e6))

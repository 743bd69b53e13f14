"""Short, fixed workload for ncu.  Usage: python tools/profile_cmd.py <config> [team] [n_envs] [presteps]

The launch schedule is pinned (DG_SPLIT=1 unless the caller sets it): under the adaptive default the first period runs fused and the
per-step launch count - what ncu's -s / -c count - would depend on when the robots touch down."""
import os
import sys

os.environ.setdefault('DG_SPLIT', '1')

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from bench import CONFIGS, ROOT, action_ranges, register_example_addons  # noqa: E402
from diy_gym_b200 import DIYGym  # noqa: E402

register_example_addons()
name = sys.argv[1] if len(sys.argv) > 1 else 'ur_high_5'
team = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
env = DIYGym(os.path.join(ROOT, CONFIGS[name][0]), num_envs=n, team=team)
w = env.world
lo, hi = action_ranges(env)
lo, hi = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
if w.n_act:
    w.action.copy_(lo + (hi - lo) * torch.rand((n, w.n_act), device='cuda'))
for _ in range(steps):
    w.step()
torch.cuda.synchronize()
print('ok', name, w.team, w.block_threads, w.grid_blocks, w.smem_bytes)

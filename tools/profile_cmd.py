"""Short, fixed workload for ncu: ur_high_5, 2048 envs, 1 reset + 3 steps.  Usage: python tools/profile_cmd.py [team]"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from diy_gym_b200.backend import World  # noqa: E402
from tools.manual_scenes import ur_high_5  # noqa: E402

team = int(sys.argv[1]) if len(sys.argv) > 1 else 0
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
w = World(ur_high_5(), n, team=team)
w.reset()
w.action.uniform_(-0.01, 0.01)
for _ in range(3):
    w.step()
torch.cuda.synchronize()
print('ok', w.team, w.block_threads, w.grid_blocks, w.smem_bytes)

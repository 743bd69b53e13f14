"""Timing sweep over team sizes for the example configs.  Usage: python tools/gpu_probe2.py [config ...]"""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from bench import CONFIGS, ROOT, action_ranges, register_example_addons  # noqa: E402
from diy_gym_b200 import DIYGym  # noqa: E402

register_example_addons()
names = sys.argv[1:] or ['ur_high_5', 'r2d2_maze', 'from_the_readme']
for name in names:
    path, n_envs = CONFIGS[name]
    for team in (1, 2, 4, 8):
        for block in (32, 64, 128):
            if block < team:
                continue
            os.environ['DG_BLOCK'] = str(block)
            try:
                env = DIYGym(os.path.join(ROOT, path), num_envs=n_envs, team=team)
            except Exception as e:
                print(name, team, block, 'ERR', str(e)[:100], flush=True)
                continue
            w = env.world
            lo, hi = action_ranges(env)
            lo, hi = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
            if w.n_act:
                w.action.copy_(lo + (hi - lo) * torch.rand((n_envs, w.n_act), device='cuda'))
            for _ in range(3):
                w.step()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(10):
                w.step()
            t1.record()
            torch.cuda.synchronize()
            ms = t0.elapsed_time(t1) / 10
            print('%-16s n %5d team %2d block %3d grid %4d smem %6d ws %5d B: %8.3f ms/step %10.0f env-steps/s' %
                  (name, n_envs, team, w.block_threads, w.grid_blocks, w.smem_bytes, w.ws_floats * 4, ms, n_envs / ms * 1e3), flush=True)
            env.close()

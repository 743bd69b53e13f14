"""One timing per (config, team, block, ws mode, envs) given on the command line: name:team:envs_per_block[:mode[:n_envs]]"""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from bench import CONFIGS, ROOT, action_ranges, register_example_addons  # noqa: E402
from diy_gym_b200 import DIYGym  # noqa: E402

register_example_addons()
for spec in sys.argv[1:]:
    parts = spec.split(':')
    name, team, block = parts[0], int(parts[1]), int(parts[2])
    if block:
        os.environ['DG_ENVS_PER_BLOCK'] = str(block)
    else:
        os.environ.pop('DG_ENVS_PER_BLOCK', None)
    os.environ['DG_WS_MODE'] = parts[3] if len(parts) > 3 else '2'
    n_envs = int(parts[4]) if len(parts) > 4 else CONFIGS[name][1]
    try:
        env = DIYGym(os.path.join(ROOT, CONFIGS[name][0]), num_envs=n_envs, team=team)
    except Exception as e:
        print(spec, 'ERR', str(e)[:120], flush=True)
        continue
    w = env.world
    lo, hi = action_ranges(env)
    lo, hi = torch.from_numpy(lo).cuda(), torch.from_numpy(hi).cuda()
    g = torch.Generator(device='cuda').manual_seed(0)
    if w.n_act:
        w.action.copy_(lo + (hi - lo) * torch.rand((n_envs, w.n_act), device='cuda', generator=g))
    for _ in range(30 if name == 'r2d2_maze' else 3):
        w.step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(10):
        w.step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 10
    print('%-22s n %5d team %2d block %3d grid %4d smem %6d ws %5d B: %8.3f ms/step %10.0f env-steps/s' %
          (spec, n_envs, team, w.block_threads, w.grid_blocks, w.smem_bytes, w.ws_floats * 4, ms, n_envs / ms * 1e3), flush=True)
    env.close()

"""First-contact GPU probe: parity of the CUDA step against the oracle on a hand-built ur_high_5 scene and a
team-size / block-size timing sweep.  Usage (GPU box): python tools/gpu_probe.py [n_envs]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from diy_gym_b200.backend import World  # noqa: E402
from oracle.oracle import OracleWorld  # noqa: E402
from tools.manual_scenes import ur_high_5  # noqa: E402


def parity(sc, team):
    n = 8
    w = World(sc, n, team=team)
    oracles = [OracleWorld(sc, env_id=i) for i in range(n)]
    w.reset()
    obs_o = np.stack([o.env_reset()[0] for o in oracles])
    torch.cuda.synchronize()
    err0 = np.abs(w.obs.cpu().numpy() - obs_o).max()
    rng = np.random.default_rng(0)
    worst = 0.0
    for k in range(20):
        a = rng.uniform(-0.01, 0.01, (n, sc['n_act']))
        w.action.copy_(torch.from_numpy(a.astype(np.float32)))
        w.step()
        outs = [o.env_step(a[i]) for i, o in enumerate(oracles)]
        torch.cuda.synchronize()
        eo = np.abs(w.obs.cpu().numpy() - np.stack([x[0] for x in outs])).max()
        er = np.abs(w.reward.cpu().numpy() - np.stack([x[1] for x in outs])).max()
        worst = max(worst, eo, er)
    print('team %2d  reset obs err %.2e  20-step rollout max obs/reward err %.2e  (block %d, grid %d, smem %d)' %
          (team, err0, worst, w.block_threads, w.grid_blocks, w.smem_bytes), flush=True)
    w.close()


def timing(sc, n_envs, team, block, steps=20):
    os.environ['DG_BLOCK'] = str(block)
    w = World(sc, n_envs, team=team)
    w.reset()
    g = torch.Generator(device='cuda').manual_seed(0)
    acts = (torch.rand((n_envs, sc['n_act']), device='cuda', generator=g) - 0.5) * 0.02
    w.action.copy_(acts)
    for _ in range(3):
        w.step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        w.step()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    print('n_envs %6d team %2d block %3d grid %4d smem %6d : %8.3f ms/step  %10.0f env-steps/s' %
          (n_envs, team, w.block_threads, w.grid_blocks, w.smem_bytes, ms, n_envs / ms * 1e3), flush=True)
    w.close()


if __name__ == '__main__':
    n_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    sc = ur_high_5()
    print(torch.cuda.get_device_name(0))
    for team in (1, 2, 4, 8, 32):
        parity(sc, team)
    for team in (1, 2, 4, 8, 16, 32):
        for block in (32, 64, 128):
            if block >= team:
                timing(sc, n_envs, team, block)
    sc8 = ur_high_5(max_contacts=8)
    for team in (2, 4, 8):
        timing(sc8, n_envs, team, 64)
    for n in (1024, 16384, 65536):
        timing(sc, n, 4, 64, steps=10)

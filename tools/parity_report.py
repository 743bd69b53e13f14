"""Rollout parity report: the CUDA path (through the C ABI) against the fp64 CPU oracle from identical initial states
and identical action sequences - short-horizon divergence of q / qdot / base pose per step, and the distribution of
episode returns (BASELINE.json north_star: "short-horizon rollout divergence and episode-return distributions must be
reported").  pybullet is absent, so the comparison arm is oracle/bullet_restatement.c (PARITY UNPINNED against pybullet).

    python tools/parity_report.py [--envs 64] [--horizon 100] [--episodes-steps 300] > profiles/<name>.json
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from bench import CONFIGS, action_ranges, register_example_addons  # noqa: E402
from diy_gym_b200 import DIYGym  # noqa: E402
from oracle.oracle import OracleWorld  # noqa: E402


def rel(a, b, floor):
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), floor)) if a.size else 0.0


def report(name, n_envs, horizon, ep_steps):
    env = DIYGym(os.path.join(ROOT, CONFIGS[name][0]), num_envs=n_envs, device=0, seed=4321)
    w, sc, h = env.world, env.scene, env.scene.hdr
    nd, nb = sc['nd'], sc['nb']
    lo, hi = action_ranges(env)
    oracles = [OracleWorld(sc, seed=4321, env_id=i) for i in range(n_envs)]
    w.reset()
    for o in oracles:
        o.env_reset()
    torch.cuda.synchronize()
    # identical initial states: the oracle's reset state (fp32-rounded) goes to both
    st = np.stack([o.state for o in oracles]).astype(np.float32)
    w.state.copy_(torch.from_numpy(st))
    w.param.copy_(torch.from_numpy(np.stack([o.param for o in oracles]).astype(np.float32)))
    for i, o in enumerate(oracles):
        o.state[:] = st[i]
        o.refresh()   # (the row carries the link cache too; refreshed anyway so that nothing depends on that)
    rng = np.random.default_rng(7)
    div = {'q': [], 'qd': [], 'base_pos': [], 'base_quat': []}
    ret_gpu, ret_cpu = np.zeros(n_envs), np.zeros(n_envs)
    term_mismatch, contacts = 0, 0
    for k in range(max(horizon, ep_steps)):
        a = rng.uniform(lo, hi, (n_envs, max(w.n_act, 1))).astype(np.float32)[:, :w.n_act]
        if w.n_act:
            w.action.copy_(torch.from_numpy(a))
        w.step()
        torch.cuda.synchronize()
        outs = [o.env_step(a[i].astype(np.float64)) for i, o in enumerate(oracles)]
        sg = w.state.cpu().numpy().astype(np.float64)
        so = np.stack([o.state for o in oracles])
        if k < horizon:
            div['q'].append(rel(sg[:, h['S_Q']:h['S_Q'] + nd], so[:, h['S_Q']:h['S_Q'] + nd], 1e-30))
            div['qd'].append(rel(sg[:, h['S_QD']:h['S_QD'] + nd], so[:, h['S_QD']:h['S_QD'] + nd], 1e-2))   # the north_star floors: max|q|, max(max|qd|, 1e-2)
            div['base_pos'].append(rel(sg[:, h['S_BPOS']:h['S_BPOS'] + 3 * nb], so[:, h['S_BPOS']:h['S_BPOS'] + 3 * nb], 1.0))
            div['base_quat'].append(rel(sg[:, h['S_BQUAT']:h['S_BQUAT'] + 4 * nb], so[:, h['S_BQUAT']:h['S_BQUAT'] + 4 * nb], 1.0))
        if k < ep_steps and w.n_rew:
            ret_gpu += w.reward.cpu().numpy().astype(np.float64).sum(axis=1)
            ret_cpu += np.stack([x[1] for x in outs]).sum(axis=1)
        if w.n_term:
            term_mismatch += int((w.term.cpu().numpy() != np.stack([x[2] for x in outs])).sum())
    pick = lambda v: {str(s): v[s - 1] for s in (1, 2, 5, 10, 20, 50, 100) if s <= len(v)}
    out = {'config': name, 'envs': n_envs, 'seed': 4321, 'actions': 'uniform in the add-on action bounds (bench.py action_ranges)',
           'max_relative_divergence_by_step': {k2: pick(v) for k2, v in div.items()},
           'terminal_flag_mismatches': term_mismatch}
    if w.n_rew:
        q = [5, 25, 50, 75, 95]
        out['episode_return_%d_steps' % ep_steps] = {
            'cuda': {'mean': float(ret_gpu.mean()), 'std': float(ret_gpu.std()), 'percentiles': dict(zip(map(str, q), np.percentile(ret_gpu, q).tolist()))},
            'oracle': {'mean': float(ret_cpu.mean()), 'std': float(ret_cpu.std()), 'percentiles': dict(zip(map(str, q), np.percentile(ret_cpu, q).tolist()))},
            'max_abs_difference_per_env': float(np.abs(ret_gpu - ret_cpu).max()),
            'max_rel_difference_per_env': float((np.abs(ret_gpu - ret_cpu) / np.maximum(np.abs(ret_cpu), 1e-9)).max())}
    return out


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--envs', type=int, default=64)
    ap.add_argument('--horizon', type=int, default=100)
    ap.add_argument('--episode-steps', type=int, default=300)
    ap.add_argument('--configs', default='ur_high_5,ur_high_5_randomised,r2d2_maze,from_the_readme,basic_env')
    args = ap.parse_args()
    register_example_addons()
    res = {'arm': 'CUDA path (fp32, C ABI) vs fp64 CPU oracle (oracle/bullet_restatement.c) - PARITY UNPINNED against pybullet (not installed)',
           'reports': [report(n, args.envs, args.horizon, args.episode_steps) for n in args.configs.split(',')]}
    print(json.dumps(res, indent=1))

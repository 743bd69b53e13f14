#!/usr/bin/env python
"""Writes tests/golden/<config>.npz: seeded rollouts of the fp64 CPU oracle on the example configs.

There is no pybullet in this container, so these vectors come from the repo's own oracle
(oracle/bullet_restatement.c) - they pin the oracle against silent change and give the CUDA path a fixed,
committed target; they are NOT pybullet outputs ("parity unpinned").

Each file holds: actions [K][n_act], and after reset and after every step the observation / reward / terminal rows and
the joint / base state, for environment ids 0 and 5 with seed 4321.
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
from bench import CONFIGS, action_ranges, register_example_addons  # noqa: E402
from diy_gym_b200 import DIYGym  # noqa: E402
from oracle.oracle import OracleWorld  # noqa: E402

K = 12
SEED = 4321
ENV_IDS = (0, 5)


def main():
    register_example_addons()
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    for name in ('ur_high_5', 'ur_high_5_randomised', 'from_the_readme', 'r2d2_maze', 'basic_env', 'ur_admittance'):
        env = DIYGym(os.path.join(ROOT, CONFIGS[name][0]), num_envs=1, compile_only=True)
        sc = env.scene
        lo, hi = action_ranges(env)
        rng = np.random.default_rng(SEED)
        actions = rng.uniform(lo, hi, (K, sc['n_act'])) if sc['n_act'] else np.zeros((K, 0))
        rec = dict(actions=actions, seed=SEED, env_ids=np.array(ENV_IDS), ibuf=sc.ibuf, fbuf_sum=np.array([sc.fbuf.sum()]))
        for eid in ENV_IDS:
            o = OracleWorld(sc, seed=SEED, env_id=eid)
            obs, rew, term = o.env_reset()
            O, R, T, ST = [obs], [rew], [term], [o.state.copy()]
            for k in range(K):
                obs, rew, term = o.env_step(actions[k])
                O.append(obs); R.append(rew); T.append(term); ST.append(o.state.copy())
            rec['obs_%d' % eid], rec['rew_%d' % eid], rec['term_%d' % eid] = np.array(O), np.array(R), np.array(T)
            rec['state_%d' % eid] = np.array(ST).astype(np.float32)
            rec['param_%d' % eid] = o.param.copy()
        path = os.path.join(out_dir, name + '.npz')
        np.savez_compressed(path, **rec)
        print('%-22s %6.1f KB  n_act %2d n_obs %3d' % (name, os.path.getsize(path) / 1024, sc['n_act'], sc['n_obs']))


if __name__ == '__main__':
    main()

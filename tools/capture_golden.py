#!/usr/bin/env python
"""Captures golden vectors from REAL pybullet: to be run wherever `import pybullet` works, with the reference
checkout on PYTHONPATH.  For every config it runs the reference's own DIYGym (unmodified; only `render: no` and, where the
box has no TinyRenderer-capable display stack, the cameras are dropped) for K random steps and records, per step,

    state_t (base pose / twist and joint q / qdot of every model), action_t, state_{t+1}, obs, reward, terminal

plus the pybullet version string into tests/golden/pybullet_<config>.npz.  tests/ can then pin the oracle - and through it
the CUDA path - against pybullet itself; until such files exist every physics parity claim in this repo reads "vs the CPU
oracle, unpinned against pybullet" (DESIGN.md section 2).

    PYTHONPATH=/path/to/diy-gym python tools/capture_golden.py [--steps 256] [--configs ur_high_5,...] [--out tests/golden]
    python tools/capture_golden.py --shim        # dry run in this repo: the oracle-backed shim stands in for pybullet
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import yaml

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def leaves(d, prefix=''):
    out = []
    for k, v in (d.items() if isinstance(d, dict) else []):
        out += leaves(v, prefix + '/' + k) if isinstance(v, dict) else [(prefix + '/' + k, np.atleast_1d(np.asarray(v, dtype=np.float64)).reshape(-1))]
    return out


def strip_cameras(node):
    for k in list(node.keys()):
        if isinstance(node[k], dict):
            if node[k].get('addon') == 'camera':
                del node[k]
            else:
                strip_cameras(node[k])


def sample(space, rng, scale):
    from gym import spaces
    if isinstance(space, spaces.Dict):
        return {k: sample(v, rng, scale) for k, v in space.spaces.items()}
    return rng.uniform(np.asarray(space.low, float) * scale, np.asarray(space.high, float) * scale)


def all_models(env):
    out = []

    def walk(models):
        for name in sorted(models):
            out.append(models[name])
            walk(models[name].models)
    walk(env.models)
    return out


def snapshot(p, models):
    s = []
    for m in models:
        pos, quat = p.getBasePositionAndOrientation(m.uid)
        lin, ang = p.getBaseVelocity(m.uid)
        s += list(pos) + list(quat) + list(lin) + list(ang)
        for j in range(p.getNumJoints(m.uid)):
            if p.getJointInfo(m.uid, j)[3] > -1:
                js = p.getJointState(m.uid, j)
                s += [js[0], js[1]]
    return np.array(s, float)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=256)
    ap.add_argument('--configs', default='ur_high_5,from_the_readme,r2d2_maze,basic_env,ur_admittance,ur_gripper')
    ap.add_argument('--out', default=os.path.join(ROOT, 'tests', 'golden'))
    ap.add_argument('--reference', default=os.environ.get('DIYGYM_REFERENCE', '/root/reference'))
    ap.add_argument('--keep-cameras', action='store_true')
    ap.add_argument('--shim', action='store_true', help='dry run on the oracle-backed shim (writes shim_<config>.npz)')
    args = ap.parse_args()
    sys.path.insert(0, ROOT)
    sys.path.insert(0, args.reference)
    if args.shim:
        sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shim'))
    import pybullet as p
    from diy_gym import DIYGym
    version = 'oracle-shim' if args.shim else str(getattr(p, 'getAPIVersion', lambda: 'unknown')())
    paths = {'ur_high_5': os.path.join(args.reference, 'examples', 'ur_high_5', 'ur_high_5.yaml'),
             'from_the_readme': os.path.join(args.reference, 'examples', 'from_the_readme', 'from_the_readme.yaml'),
             'basic_env': os.path.join(args.reference, 'diy_gym', 'tests', 'basic_env.yaml'),
             'r2d2_maze': os.path.join(ROOT, 'examples', 'r2d2_maze', 'r2d2_maze.yaml'),
             'ur_admittance': os.path.join(ROOT, 'examples', 'ur_admittance', 'ur_admittance.yaml'),
             'ur_gripper': os.path.join(ROOT, 'examples', 'ur_gripper', 'ur_gripper.yaml')}
    for name in args.configs.split(','):
        node = yaml.load(open(paths[name]), Loader=yaml.FullLoader)
        node['render'] = False
        if not args.keep_cameras:
            strip_cameras(node)
        tmp = os.path.join(tempfile.mkdtemp(), name + '.yaml')
        yaml.dump(node, open(tmp, 'w'))
        np.random.seed(0)
        env = DIYGym(tmp)
        models = all_models(env)
        rng = np.random.default_rng(4321)
        scale = 20.0 if name == 'r2d2_maze' else 1.0
        obs = env.reset()
        rec = {'pybullet_version': np.array([version]), 'model_names': np.array([m.name for m in models]),
               'obs_keys': np.array([k for k, _ in leaves(obs)]), 'obs_0': np.concatenate([v for _, v in leaves(obs)] or [np.zeros(0)]),
               'state_0': snapshot(p, models)}
        for k in range(1, args.steps + 1):
            action = sample(env.action_space, rng, scale)
            al = leaves(action)
            rec['act_keys'] = np.array([kk for kk, _ in al])
            rec['act_%d' % k] = np.concatenate([v for _, v in al] or [np.zeros(0)])
            obs, rew, term, _ = env.step(action)
            rec['state_%d' % k] = snapshot(p, models)
            rec['obs_%d' % k] = np.concatenate([v for _, v in leaves(obs)] or [np.zeros(0)])
            rec['rew_%d' % k] = np.concatenate([v for _, v in leaves(rew)] or [np.zeros(0)]) if isinstance(rew, dict) else np.atleast_1d(float(rew))
            rec['term_%d' % k] = np.concatenate([v for _, v in leaves(term)] or [np.zeros(0)]) if isinstance(term, dict) else np.atleast_1d(float(term))
        dst = os.path.join(args.out, ('shim_' if args.shim else 'pybullet_') + name + '.npz')
        np.savez_compressed(dst, **rec)
        print('%-16s %d steps, state %d floats, obs %d floats, pybullet %s -> %s' % (name, args.steps, rec['state_0'].size, rec['obs_0'].size, version, dst))
        env.close()


if __name__ == '__main__':
    main()

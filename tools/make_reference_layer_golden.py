#!/usr/bin/env python
"""Runs the REFERENCE's own Python layer (/root/reference/diy_gym: DIYGym, Model, Configuration, every add-on,
unmodified, imported in this container only) on top of the oracle-backed `pybullet` shim (oracle/shim) and records
what it returns -> tests/golden/reflayer_<config>.npz.

What this pins: the add-on arithmetic and bookkeeping of the reference (joint selection and ordering, IK call and
slicing, motor gains, sensor frames and quaternion conventions, reward / terminal formulas, nested dict structure and
ordering, episode timer) as executed by the reference's code.  What it does not pin: the physics below the pybullet
API, which here is this repo's oracle (pybullet itself is absent).
Cameras are stripped from the configs (the shim does not implement getCameraImage); configs are otherwise the
reference's own files plus `render: no`.
"""
import importlib.util
import os
import sys
import tempfile

import numpy as np
import yaml

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'shim'))

import pybullet as p  # noqa: E402  (the shim)
from diy_gym import DIYGym  # noqa: E402  (the reference)

K = 10
CONFIGS = {
    'ur_high_5': os.path.join(REF, 'examples', 'ur_high_5', 'ur_high_5.yaml'),
    'from_the_readme': os.path.join(REF, 'examples', 'from_the_readme', 'from_the_readme.yaml'),
    'drone_pilot': os.path.join(REF, 'examples', 'drone_pilot', 'drone_pilot.yaml'),
    'basic_env': os.path.join(REF, 'diy_gym', 'tests', 'basic_env.yaml'),
    'r2d2_maze': os.path.join(ROOT, 'examples', 'r2d2_maze', 'r2d2_maze.yaml'),   # emitted by the reference's generator
    # the reference ships no config for these add-ons / nested models; these two use its schema and its own classes
    'ur_admittance': os.path.join(ROOT, 'examples', 'ur_admittance', 'ur_admittance.yaml'),
    'ur_gripper': os.path.join(ROOT, 'examples', 'ur_gripper', 'ur_gripper.yaml'),
    'ur_extras': os.path.join(ROOT, 'examples', 'ur_extras', 'ur_extras.yaml'),
    'ur_robotiq': os.path.join(ROOT, 'examples', 'ur_gripper', 'ur_robotiq.yaml'),
}


def strip_cameras(node):
    for k in list(node.keys()):
        v = node[k]
        if isinstance(v, dict):
            if v.get('addon') == 'camera':
                del node[k]
            else:
                strip_cameras(v)


def leaves(d, prefix=''):
    out = []
    for k, v in d.items():
        if isinstance(v, dict):
            out += leaves(v, prefix + '/' + k)
        else:
            out.append((prefix + '/' + k, np.atleast_1d(np.asarray(v, dtype=np.float64)).reshape(-1)))
    return out


def sample_action(space, rng, scale):
    from gym import spaces
    if isinstance(space, spaces.Dict):
        return {k: sample_action(v, rng, scale) for k, v in space.spaces.items()}
    lo, hi = np.asarray(space.low, float) * scale, np.asarray(space.high, float) * scale
    return rng.uniform(lo, hi)


def main():
    # the drone example registers its two user add-ons at import time
    spec = importlib.util.spec_from_file_location('ref_drone_pilot', os.path.join(REF, 'examples', 'drone_pilot', 'drone_pilot.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    for name, path in CONFIGS.items():
        node = yaml.load(open(path), Loader=yaml.FullLoader)
        node['render'] = False
        strip_cameras(node)
        tmp = os.path.join(tempfile.mkdtemp(), name + '.yaml')
        yaml.dump(node, open(tmp, 'w'))
        np.random.seed(0)
        env = DIYGym(tmp)
        rng = np.random.default_rng(123)
        scale = 20.0 if name == 'r2d2_maze' else 1.0
        rec = {}
        obs = env.reset()
        pose0 = {n: np.r_[p.getBasePositionAndOrientation(m.uid)[0], p.getBasePositionAndOrientation(m.uid)[1]] for n, m in env.models.items()}
        rec['obs_keys'] = np.array([k for k, _ in leaves(obs)])
        rec['obs_0'] = np.concatenate([v for _, v in leaves(obs)]) if len(obs) else np.zeros(0)
        rec['pose0_names'] = np.array(sorted(pose0))
        rec['pose0'] = np.array([pose0[n] for n in sorted(pose0)])
        act_keys = None
        for k in range(K):
            action = sample_action(env.action_space, rng, scale)
            al = leaves(action)
            act_keys = [kk for kk, _ in al]
            rec['act_%d' % (k + 1)] = np.concatenate([v for _, v in al]) if al else np.zeros(0)
            obs, rew, term, _ = env.step(action)
            rec['obs_%d' % (k + 1)] = np.concatenate([v for _, v in leaves(obs)]) if len(obs) else np.zeros(0)
            if isinstance(rew, dict):
                rl = leaves(rew)
                rec['rew_keys'] = np.array([kk for kk, _ in rl])
                rec['rew_%d' % (k + 1)] = np.concatenate([v for _, v in rl]) if rl else np.zeros(0)
            else:
                rec['rew_keys'] = np.array(['<collapsed>'])
                rec['rew_%d' % (k + 1)] = np.atleast_1d(float(rew))
            if isinstance(term, dict):
                tl = leaves(term)
                rec['term_keys'] = np.array([kk for kk, _ in tl])
                rec['term_%d' % (k + 1)] = np.concatenate([v for _, v in tl]) if tl else np.zeros(0)
            else:
                rec['term_keys'] = np.array(['<collapsed>'])
                rec['term_%d' % (k + 1)] = np.atleast_1d(float(term))
        poseK = {n: np.r_[p.getBasePositionAndOrientation(m.uid)[0], p.getBasePositionAndOrientation(m.uid)[1]] for n, m in env.models.items()}
        rec['poseK'] = np.array([poseK[n] for n in sorted(poseK)])
        rec['act_keys'] = np.array(act_keys if act_keys else [])
        rec['steps'] = np.array([K])
        dst = os.path.join(out_dir, 'reflayer_' + name + '.npz')
        np.savez_compressed(dst, **rec)
        print('%-18s obs %4d floats %s  rewards %s  terminals %s  -> %.1f KB' % (name, rec['obs_0'].size, list(rec['obs_keys'])[:3], list(rec['rew_keys']), list(rec['term_keys']), os.path.getsize(dst) / 1024))
        env.close()


if __name__ == '__main__':
    main()

"""This repo's DIYGym (host layer + fused add-on ops) against records of the REFERENCE's own Python layer
(tests/golden/reflayer_*.npz, written by tools/make_reference_layer_golden.py: /root/reference/diy_gym unmodified,
running on the oracle-backed pybullet shim).  Same actions in, same nested keys in the same order and the same
observation / reward / terminal values out.  The physics under both is this repo's engine (oracle there, fp32 kernel
code here), so the tolerance is the fp32-vs-fp64 rollout tolerance; pybullet itself is absent.
CPU: kernel code built by tests/emul.  GPU (-m gpu): the CUDA path."""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch
import yaml

from bench import CONFIGS, register_example_addons
from diy_gym_b200 import Configuration, DIYGym

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
NAMES = ['ur_high_5', 'from_the_readme', 'drone_pilot', 'basic_env', 'r2d2_maze', 'ur_admittance', 'ur_gripper', 'ur_extras', 'ur_robotiq']
# ur_gripper: the welded child's spawn transient is solved with clamped, unconverged sweeps whose result depends on
# rounding; the x86 build of the kernel code follows the fp64 rollout from reset, the GPU build is
# compared after the transient in tests/test_gpu_parity.py instead
GPU_NAMES = [n for n in NAMES if n not in ('ur_gripper', 'ur_robotiq')]
TOL = dict(rtol=3e-3, atol=3e-4)
# ur_robotiq: as in the reference (model.py:69-77) the gripper is spawned AT the parent frame, overlapping the wrist it is then welded
# to 1 cm away; hull contacts fight the weld for the first steps and the light wrist joints are thrown about at up to 25 rad/s in a
# way that depends on rounding (both arms: the oracle as well).  Compared there: nested keys and their order, terminals, and the three
# heavy joints (shoulder pan / lift, elbow) to 2e-2 rad
LOOSE = {'ur_robotiq': dict(rtol=0.0, atol=2e-2)}


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
            out.append((prefix + '/' + k, v))
    return out


def flat(d):
    ls = leaves(d)
    return [k for k, _ in ls], (np.concatenate([np.asarray(v.detach().cpu() if isinstance(v, torch.Tensor) else v, dtype=np.float64).reshape(-1) for _, v in ls]) if ls else np.zeros(0))


def build_action(keys, values, space):
    """golden flat action (reference leaf order) -> nested dict of [1, n] tensors with this repo's space shapes."""
    action, pos = OrderedDict(), 0
    for key in keys:
        parts = key.strip('/').split('/')
        sp = space
        for pth in parts:
            sp = sp.spaces[pth]
        n = int(np.prod(sp.shape))
        node = action
        for pth in parts[:-1]:
            node = node.setdefault(pth, OrderedDict())
        node[parts[-1]] = torch.as_tensor(values[pos:pos + n], dtype=torch.float32).reshape(1, *sp.shape)
        pos += n
    return action


def make_env(name, g, world_factory=None, **kw):
    register_example_addons()
    node = yaml.load(open(os.path.join(ROOT, CONFIGS[name][0])), Loader=yaml.FullLoader)
    strip_cameras(node)
    if name == 'drone_pilot':   # the reference draws the target pose from the global numpy RNG: replay its draw
        names = list(g['pose0_names'])
        node['target']['xyz'] = [float(x) for x in g['pose0'][names.index('target')][:3]]
        node['target']['respawn']['position_range'] = [0.0, 0.0, 0.0]
    return DIYGym(Configuration.from_dict(name, node), num_envs=1, world_factory=world_factory, **kw)


def check(name, env):
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'reflayer_' + name + '.npz'))
    obs = env.reset()
    keys, vals = flat(obs)
    assert keys == list(g['obs_keys'])                       # same nested structure, same order
    tol = LOOSE.get(name, TOL)
    pos_only = (lambda v: v[:3]) if name in LOOSE else (lambda v: v)
    assert np.allclose(pos_only(vals), pos_only(g['obs_0']), **tol)
    for k in range(1, int(g['steps'][0]) + 1):
        action = build_action(list(g['act_keys']), g['act_%d' % k], env.action_space)
        dev = env.world.state.device
        action = OrderedDict((r, OrderedDict((a, (OrderedDict((kk, vv.to(dev)) for kk, vv in v.items()) if isinstance(v, dict) else v.to(dev))) for a, v in x.items()))
                             for r, x in action.items())
        obs, rew, term, _ = env.step(action)
        keys, vals = flat(obs)
        assert keys == list(g['obs_keys'])
        assert np.allclose(pos_only(vals), pos_only(g['obs_%d' % k]), **tol), (name, k, np.abs(vals - g['obs_%d' % k]).max())
        if isinstance(rew, dict):
            rk, rv = flat(rew)
            assert rk == list(g['rew_keys'])
        else:
            assert list(g['rew_keys']) == ['<collapsed>']
            rv = np.atleast_1d(rew.detach().cpu().numpy().astype(np.float64))
        assert np.allclose(rv, g['rew_%d' % k], rtol=5e-3, atol=5e-4), (name, k)
        if isinstance(term, dict):
            tk, tv = flat(term)
            assert tk == list(g['term_keys'])
        else:
            assert list(g['term_keys']) == ['<collapsed>']
            tv = np.atleast_1d(term.detach().cpu().numpy().astype(np.float64))
        assert np.array_equal(tv, g['term_%d' % k]), (name, k)
    names = sorted(env.models)
    assert names == list(g['pose0_names'])
    mine = np.array([np.r_[env.models[n].base_pose()[0][0].cpu().numpy(), env.models[n].base_pose()[1][0].cpu().numpy()] for n in names])
    if name not in LOOSE:
        assert np.allclose(mine, g['poseK'], **TOL)


@pytest.mark.parametrize('name', NAMES)
def test_host_layer_matches_reference_layer_cpu(name):
    from tests.emul.world import factory
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'reflayer_' + name + '.npz'))
    check(name, make_env(name, g, world_factory=factory(team=8 if name in ('ur_gripper', 'ur_robotiq') else 4)))


@pytest.mark.gpu
@pytest.mark.parametrize('name', GPU_NAMES)
def test_host_layer_matches_reference_layer_cuda(name):
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'reflayer_' + name + '.npz'))
    env = make_env(name, g, device=0)
    check(name, env)
    env.close()

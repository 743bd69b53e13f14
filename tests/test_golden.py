"""REGRESSION FIXTURES, not parity evidence: committed rollouts of the repo's OWN fp64 oracle (tests/golden/*.npz, written by
tools/make_golden.py and regenerated whenever the oracle changes on purpose - pybullet is absent).
They pin the oracle and the compiler against silent change and give the CUDA path a fixed target that travels to the GPU box.  CPU: the oracle and the g++ build of the kernel code reproduce them; GPU (-m gpu):
the CUDA path reproduces them through the C ABI."""
import os

import numpy as np
import pytest

from bench import CONFIGS, register_example_addons
from diy_gym_b200 import DIYGym
from oracle.oracle import OracleWorld

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
NAMES = ['ur_high_5', 'ur_high_5_randomised', 'from_the_readme', 'r2d2_maze', 'basic_env', 'ur_admittance']
# open-loop rollout of K = 12 steps: fp32 vs fp64 round-off grows a little along the rollout
TOL = dict(rtol=2e-3, atol=2e-4)


def load(name):
    g = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
    register_example_addons()
    env = DIYGym(os.path.join(ROOT, CONFIGS[name][0]), num_envs=1, compile_only=True)
    assert np.array_equal(env.scene.ibuf, g['ibuf']), 'scene compiler output changed: regenerate with tools/make_golden.py'
    return g, env.scene


@pytest.mark.parametrize('name', NAMES)
def test_oracle_reproduces_golden(name):
    g, sc = load(name)
    for eid in g['env_ids']:
        o = OracleWorld(sc, seed=int(g['seed']), env_id=int(eid))
        obs, rew, term = o.env_reset()
        assert np.allclose(obs, g['obs_%d' % eid][0], rtol=1e-10, atol=1e-12)
        for k, a in enumerate(g['actions']):
            obs, rew, term = o.env_step(a)
            assert np.allclose(obs, g['obs_%d' % eid][k + 1], rtol=1e-9, atol=1e-11)
            assert np.allclose(rew, g['rew_%d' % eid][k + 1], rtol=1e-9, atol=1e-11)
            assert np.array_equal(term, g['term_%d' % eid][k + 1])


def _check_world_against_golden(world_states_fn, g, sc, step_fn, reset_fn):
    nd, nb = sc['nd'], sc['nb']
    h = sc.hdr
    obs, rew, term = reset_fn()
    for i, eid in enumerate(g['env_ids']):
        assert np.allclose(obs[i], g['obs_%d' % eid][0], **TOL)
    for k, a in enumerate(g['actions']):
        obs, rew, term = step_fn(np.stack([a, a]))
        st = world_states_fn()
        for i, eid in enumerate(g['env_ids']):
            ref = g['state_%d' % eid][k + 1]
            for key, n in (('S_Q', nd), ('S_QD', nd), ('S_BPOS', 3 * nb), ('S_BQUAT', 4 * nb)):
                if n:
                    scale = 10.0 if key == 'S_QD' else 1.0
                    assert np.allclose(st[i, h[key]:h[key] + n], ref[h[key]:h[key] + n], rtol=TOL['rtol'] * scale, atol=TOL['atol'] * scale), (key, k)
            assert np.allclose(obs[i], g['obs_%d' % eid][k + 1], **TOL), k
            assert np.allclose(rew[i], g['rew_%d' % eid][k + 1], rtol=5e-3, atol=5e-4), k
            assert np.array_equal(term[i], g['term_%d' % eid][k + 1]), k


@pytest.mark.parametrize('name', NAMES)
def test_kernel_code_on_cpu_reproduces_golden(name):
    from tests.emul.emul import EmulWorld
    g, sc = load(name)
    # two single-environment worlds so that each gets its own global env id (0 and 5)
    team = 8 if sc['ncons'] else 4   # (welded models need the row capacity of a team of 8)
    ws = [EmulWorld(sc, 1, team, seed=int(g['seed']), env_off=int(e)) for e in g['env_ids']]
    cat = lambda outs: tuple(np.concatenate([o[j] for o in outs]) for j in range(3))
    _check_world_against_golden(lambda: np.concatenate([w.state for w in ws]), g, sc,
                                lambda a: cat([w.step(a[i:i + 1]) for i, w in enumerate(ws)]), lambda: cat([w.reset() for w in ws]))


@pytest.mark.gpu
@pytest.mark.parametrize('name', NAMES)
def test_cuda_path_reproduces_golden(name):
    import torch
    from diy_gym_b200.backend import World
    g, sc = load(name)
    ws = [World(sc, 1, seed=int(g['seed']), env_id_offset=int(e)) for e in g['env_ids']]

    def outs():
        torch.cuda.synchronize()
        return tuple(np.concatenate([getattr(w, k).cpu().numpy() for w in ws]) for k in ('obs', 'reward', 'term'))

    def step(a):
        for i, w in enumerate(ws):
            if w.n_act:
                w.action.copy_(torch.from_numpy(a[i:i + 1].astype(np.float32)))
            w.step()
        return outs()

    def reset():
        for w in ws:
            w.reset()
        return outs()
    _check_world_against_golden(lambda: np.concatenate([w.state.cpu().numpy() for w in ws]), g, sc, step, reset)
    for w in ws:
        w.close()

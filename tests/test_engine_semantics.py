"""The engine semantics SURVEY Appendix A could only recall are NAMED SWITCHES of the scene header (compiler/scene.py SEMANTICS,
extension key `engine_semantics`), honoured by the kernels and by the oracle alike: the day a pybullet golden vector disagrees with
a default, the fix is a flag (tests/test_pybullet_golden.py consumes such vectors).  Each switch is flipped here: the kernel source
(g++ build on CPU, CUDA on the GPU) stays on the oracle, and the outcome moves the way the switch says."""
import os

import numpy as np
import pytest
import torch
import yaml

from bench import CONFIGS
from diy_gym_b200 import Configuration, DIYGym
from oracle.oracle import OracleWorld

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')


def _env(name, sem, factory, n=2):
    node = yaml.load(open(os.path.join(ROOT, CONFIGS[name][0])), Loader=yaml.FullLoader)
    node['engine_semantics'] = list(sem)
    for k in [k for k, v in node.items() if isinstance(v, dict) and v.get('addon') == 'camera']:
        del node[k]
    return DIYGym(Configuration.from_dict(name, node), num_envs=n, device=0, world_factory=factory, seed=3)


def _rollout(name, sem, factory, action, steps=3):
    env = _env(name, sem, factory)
    w, sc = env.world, env.scene
    o = OracleWorld(sc, seed=3, env_id=0)
    o.env_reset()
    a = np.asarray(action, np.float32)
    for _ in range(steps):
        w.action[:] = torch.from_numpy(a).to(w.action.device)
        w.step()
        o.env_step(a.astype(np.float64))
    st = w.state[0].cpu().numpy().astype(np.float64)
    env.close()
    return st, o.state.copy(), sc


def _check_flag(factory):
    # wrench_first_substep: a marble pushed by external_force gains half the velocity per step
    push = [10.0, 0, 0, 0, 0, 0]
    base, ob, sc = _rollout('basic_env', (), factory, push, steps=1)
    half, oh, _ = _rollout('basic_env', ('wrench_first_substep', ), factory, push, steps=1)
    h = sc.hdr
    names = [b.name for b in sc.bodies]
    vb = h['S_BVEL'] + 3 * names.index('blue_marble')
    assert base[vb] > 1e-3 and abs(half[vb] / base[vb] - 0.5) < 0.05
    assert np.allclose(half, oh, rtol=2e-3, atol=2e-4) and np.allclose(base, ob, rtol=2e-3, atol=2e-4)
    # motor_clamp_substep: a saturated position motor (a 1 rad target jump on the UR5) delivers half the impulse per sub-step
    jump = [1.0, -0.73, -1.93, -0.36, -0.03, -0.06]
    base, ob, sc = _rollout('ur_extras', (), factory, jump, steps=1)
    sub, os_, _ = _rollout('ur_extras', ('motor_clamp_substep', ), factory, jump, steps=1)
    h = sc.hdr
    ia = h['S_MAPPLIED']
    assert abs(base[ia]) > 1e-3 and abs(sub[ia] / base[ia] - 0.5) < 0.02
    assert np.allclose(sub, os_, rtol=2e-3, atol=2e-4)
    # damping_linear: a drone falling at ~1 m/s is damped by k m v instead of k m v (1 + |v|)
    idle = [0.0, 0.0, 0.0, 0.0]
    base, ob, sc = _rollout('drone_pilot', (), factory, idle, steps=40)
    lin, ol, _ = _rollout('drone_pilot', ('damping_linear', ), factory, idle, steps=40)
    h = sc.hdr
    names = [b.name for b in sc.bodies]
    vz = h['S_BVEL'] + 3 * names.index('drone') + 2
    assert base[vz] < -0.5 and lin[vz] < base[vz] - 1e-4          # falls faster with the weaker damping
    assert np.allclose(lin, ol, rtol=2e-3, atol=2e-4)


def test_switches_move_kernel_and_oracle_together_cpu():
    from bench import register_example_addons
    from tests.emul.world import factory
    register_example_addons()
    _check_flag(factory(4))


@pytest.mark.gpu
def test_switches_move_kernel_and_oracle_together_gpu():
    from bench import register_example_addons
    register_example_addons()
    _check_flag(None)

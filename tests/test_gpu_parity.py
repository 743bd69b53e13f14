"""GPU parity tests proper: the CUDA path (through the C ABI, diy_gym_b200.backend.World) against the fp64 CPU
oracle on identical seeded inputs.  Tolerances follow BASELINE.json north_star: single-step physics state within
1e-4 relative (contact-free), observations / rewards within 1e-5 relative given identical states (fp32 kernel
vs fp64 oracle, so an absolute floor of a few fp32 ulps of the quantities involved is allowed)."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')
pytestmark = pytest.mark.gpu

from diy_gym_b200.assets import resolve_model  # noqa: E402
from diy_gym_b200.compiler.scene import SceneBuilder  # noqa: E402
from oracle.oracle import OracleWorld  # noqa: E402
from tools.manual_scenes import ur_high_5  # noqa: E402


def _world(sc, n, **kw):
    from diy_gym_b200.backend import World
    return World(sc, n, **kw)


def _sync_state(w, oracles):
    st = np.stack([o.state for o in oracles]).astype(np.float32)
    pr = np.stack([o.param for o in oracles]).astype(np.float32)
    w.state.copy_(torch.from_numpy(st))
    w.param.copy_(torch.from_numpy(pr))


@pytest.mark.parametrize('team', [1, 2, 4, 8, 16, 32])
def test_ur_high_5_single_step_from_identical_states(team):
    sc = ur_high_5()
    n, nd = 16, sc['nd']
    w = _world(sc, n, team=team)
    oracles = [OracleWorld(sc, env_id=i) for i in range(n)]
    rng = np.random.default_rng(team)
    for o in oracles:
        o.env_reset()
        for _ in range(int(rng.integers(0, 5))):   # decorrelate the environments
            o.env_step(rng.uniform(-0.01, 0.01, sc['n_act']))
    for k in range(6):
        _sync_state(w, oracles)
        a = rng.uniform(-0.01, 0.01, (n, sc['n_act']))
        w.action.copy_(torch.from_numpy(a.astype(np.float32)))
        w.step()
        outs = [o.env_step(a[i]) for i, o in enumerate(oracles)]
        torch.cuda.synchronize()
        q_o = np.stack([o.s('S_Q', nd) for o in oracles])
        qd_o = np.stack([o.s('S_QD', nd) for o in oracles])
        q_g = w.s('S_Q', nd).cpu().numpy()
        qd_g = w.s('S_QD', nd).cpu().numpy()
        assert np.abs(q_g - q_o).max() <= 1e-4 * np.abs(q_o).max()
        assert np.abs(qd_g - qd_o).max() <= 1e-4 * max(np.abs(qd_o).max(), 1e-2)
        obs_o = np.stack([x[0] for x in outs])
        rew_o = np.stack([x[1] for x in outs])
        # sensors read the state the step produced, so the state tolerance carries over
        assert np.allclose(w.obs.cpu().numpy(), obs_o, rtol=1e-4, atol=2e-5)
        assert np.allclose(w.reward.cpu().numpy(), rew_o, rtol=1e-4, atol=2e-5)
        assert np.array_equal(w.term.cpu().numpy(), np.stack([x[2] for x in outs]))
    w.close()


def test_observations_match_given_identical_state():
    """Add-on arithmetic only (no physics): observe after a reset-from-synced state, 1e-5 relative."""
    sc = ur_high_5()
    n = 8
    w = _world(sc, n)
    oracles = [OracleWorld(sc, env_id=i) for i in range(n)]
    w.reset()
    obs_o = np.stack([o.env_reset()[0] for o in oracles])
    rew_o = np.stack([o.observe()[1] for o in oracles])
    torch.cuda.synchronize()
    assert np.allclose(w.obs.cpu().numpy(), obs_o, rtol=1e-5, atol=1e-6)
    assert np.allclose(w.reward.cpu().numpy(), rew_o, rtol=1e-5, atol=1e-6)
    w.close()


def test_rollout_divergence_is_bounded_contact_free():
    sc = ur_high_5()
    n = 8
    w = _world(sc, n)
    oracles = [OracleWorld(sc, env_id=i) for i in range(n)]
    w.reset()
    for o in oracles:
        o.env_reset()
    rng = np.random.default_rng(3)
    worst = 0.0
    for k in range(100):
        a = rng.uniform(-0.01, 0.01, (n, sc['n_act']))
        w.action.copy_(torch.from_numpy(a.astype(np.float32)))
        w.step()
        outs = [o.env_step(a[i]) for i, o in enumerate(oracles)]
    torch.cuda.synchronize()
    worst = np.abs(w.obs.cpu().numpy() - np.stack([x[0] for x in outs])).max()
    assert worst < 5e-3, worst   # 100-step open-loop divergence of an fp32 vs fp64 rollout


def test_free_body_and_contacts_drone_lands_on_plane():
    sb = SceneBuilder()
    sb.add_body('plane', resolve_model('grass/plane.urdf'))
    sb.add_body('drone', resolve_model('hector_quadrotor/quadrotor.urdf'), xyz=(0, 0, 1.0), mass=4.0)
    sc = sb.finalize()
    w = _world(sc, 4, team=4)
    o = OracleWorld(sc)
    o.s('S_BOMEGA', 6)[3:] = [0.5, -0.3, 1.0]
    o.s('S_BVEL', 6)[3:] = [1, 0.5, 0]
    o.refresh()
    _sync_state(w, [o] * 4)
    for k in range(60):
        w.step()
        o.step_physics()
    torch.cuda.synchronize()
    for name, nn in (('S_BPOS', 6), ('S_BQUAT', 8), ('S_BVEL', 6), ('S_BOMEGA', 6)):
        assert np.allclose(w.s(name, nn).cpu().numpy()[0], o.s(name, nn), rtol=1e-4, atol=1e-4), name
    for k in range(400):
        w.step()
        o.step_physics()
    torch.cuda.synchronize()
    assert len(o.contacts()) >= 3
    assert abs(w.s('S_BPOS', 6).cpu().numpy()[2, 5] - o.s('S_BPOS', 6)[5]) < 1e-3
    assert np.abs(w.s('S_BVEL', 6).cpu().numpy()[:, 3:]).max() < 1e-3
    w.close()


def test_masked_reset_only_touches_masked_environments():
    sc = ur_high_5()
    w = _world(sc, 6)
    w.reset()
    w.action.uniform_(-0.01, 0.01)
    for _ in range(5):
        w.step()
    torch.cuda.synchronize()
    before = w.state.clone()
    mask = torch.tensor([1, 0, 0, 1, 0, 0], dtype=torch.uint8)
    w.reset(mask)
    torch.cuda.synchronize()
    after = w.state
    assert torch.equal(before[[1, 2, 4, 5]], after[[1, 2, 4, 5]])
    assert not torch.equal(before[[0, 3]], after[[0, 3]])
    assert float(after[0, sc.hdr['S_STEP']]) == 0.0
    w.close()


def test_library_fails_loudly_without_buffers_or_bad_scene():
    import ctypes
    from diy_gym_b200.backend import load_library
    L = load_library()
    h = ctypes.c_void_p()
    bad = np.zeros(64, np.int32)
    rc = L.dg_world_create(bad.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), 64, np.zeros(4).ctypes.data_as(ctypes.POINTER(ctypes.c_double)), 4, 1, 0, 0, ctypes.byref(h))
    assert rc == -2 and b'scene' in L.dg_last_error(None)

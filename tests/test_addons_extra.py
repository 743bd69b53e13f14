"""Add-on ops and entry points that no other test touched ("what's missing" 7 and "what's weak" 2, 3):
stuck_joint_cost, dg_observe (add-on arithmetic at 1e-5 relative from identical states, every config, no physics step),
dg_step_host (bit-equal to dg_step + manual copies), the contact-capacity counter, auto_reset returning the terminal step's reward.
CPU legs run the g++ build of the kernel source (tests/emul); `gpu` legs the CUDA path through the C ABI."""
import os

import numpy as np
import pytest
import torch

from bench import CONFIGS, register_example_addons
from diy_gym_b200 import Configuration, DIYGym
from oracle.oracle import OracleWorld

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
ALL = ['ur_high_5', 'ur_high_5_randomised', 'from_the_readme', 'r2d2_maze', 'basic_env', 'ur_admittance', 'ur_gripper', 'ur_extras', 'drone_pilot']


def _factory(team=4):
    from tests.emul.world import factory
    return factory(team)


def _env(name, n, factory=None, **kw):
    register_example_addons()
    return DIYGym(os.path.join(ROOT, CONFIGS[name][0]), num_envs=n, device=0, world_factory=factory, **kw)


# ---------------------------------------------------------------- stuck_joint_cost ------------------------------------
STUCK = {'ur5': {'model': 'ur5/ur5_robot.urdf', 'controller': {'addon': 'joint_controller', 'control_mode': 'position'},
                 'stuck': {'addon': 'stuck_joint_cost', 'multiplier': 0.7}}}


def _stuck(factory):
    env = DIYGym(Configuration.from_dict('stuck', STUCK), num_envs=3, device=0, world_factory=factory)
    w, sc, h = env.world, env.scene, env.scene.hdr
    lim = sc.sec['LINK_F'][:, 20:22]
    dof = sc.sec['LINK_I'][:, 3]
    lo = np.array([lim[i, 0] for i in range(len(dof)) if dof[i] >= 0])
    q = np.zeros((3, sc['nd']), np.float32)
    q[1, 2] = lo[2] + 0.005                  # environment 1: elbow 5 mrad from its lower limit -> stuck
    q[2, 2] = lo[2] + 0.05                   # environment 2: 50 mrad away -> not stuck
    w.state[:, h['S_Q']:h['S_Q'] + sc['nd']] = torch.from_numpy(q).to(w.state.device)
    w.observe()
    rew = env.reward()['ur5']['stuck'].cpu().numpy()
    assert np.allclose(rew, [0.0, -0.7, 0.0])
    # and the oracle says the same (rewards/stuck_joint_cost.py intent: the reference raises NameError)
    for i in range(3):
        o = OracleWorld(sc, env_id=i)
        o.state[:] = w.state[i].cpu().numpy()
        o.refresh()
        assert np.allclose(o.observe()[1], w.reward[i].cpu().numpy(), rtol=1e-6, atol=1e-7)
    env.close()


def test_stuck_joint_cost_cpu():
    _stuck(_factory())


@pytest.mark.gpu
def test_stuck_joint_cost_gpu():
    _stuck(None)


# ---------------------------------------------------------------- observe(): 1e-5 from identical states ----------------
def _observe_parity(name, factory, n=8):
    env = _env(name, n, factory)
    w, sc = env.world, env.scene
    oracles = [OracleWorld(sc, seed=1234, env_id=i) for i in range(n)]
    rng = np.random.default_rng(3)
    from bench import action_ranges
    lo, hi = action_ranges(env)
    for i, o in enumerate(oracles):
        o.env_reset()
        for _ in range(3 + i):
            o.env_step(rng.uniform(lo, hi))
    st = np.stack([o.state for o in oracles]).astype(np.float32)
    w.state.copy_(torch.from_numpy(st))
    w.param.copy_(torch.from_numpy(np.stack([o.param for o in oracles]).astype(np.float32)))
    w.observe()                                           # link cache + sensors / rewards / terminals, no physics
    outs = []
    for i, o in enumerate(oracles):
        o.state[:] = st[i]
        o.refresh()
        outs.append(o.observe())
    obs_o, rew_o, term_o = [np.stack([x[k] for x in outs]) for k in range(3)]
    # north_star: "observations and rewards must match within 1e-5 relative given identical states" - relative to the size of the
    # quantities an entry is computed from (positions of ~1 m, angles of ~pi): an entry that is a small difference of two of them
    # carries their fp32 round-off
    scale_o = max(1.0, float(np.abs(obs_o).max())) if obs_o.size else 1.0
    scale_r = max(1.0, float(np.abs(rew_o).max())) if rew_o.size else 1.0
    assert np.abs(w.obs.cpu().numpy() - obs_o).max(initial=0) <= 1e-5 * scale_o, name
    assert np.abs(w.reward.cpu().numpy() - rew_o).max(initial=0) <= 1e-5 * scale_r, name
    assert np.array_equal(w.term.cpu().numpy(), term_o)
    env.close()


@pytest.mark.parametrize('name', ALL)
def test_observe_matches_oracle_from_identical_state_cpu(name):
    _observe_parity(name, _factory(8 if name == 'ur_gripper' else 4))


@pytest.mark.gpu
@pytest.mark.parametrize('name', ALL)
def test_observe_matches_oracle_from_identical_state_gpu(name):
    _observe_parity(name, None, n=32)


# ---------------------------------------------------------------- dg_step_host -------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('name', ['ur_high_5', 'r2d2_maze', 'from_the_readme'])
def test_step_host_is_step_plus_copies(name):
    """The host-buffer entry point (the one the e2e number runs through) against dg_step with the copies done by hand: bit-equal."""
    n = 64
    a, b = _env(name, n, seed=9), _env(name, n, seed=9)
    wa, wb = a.world, b.world
    rng = np.random.default_rng(0)
    from bench import action_ranges
    lo, hi = action_ranges(a)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
    obs_h, rew_h, term_h = pin((n, max(wa.n_obs, 1)), torch.float32), pin((n, max(wa.n_rew, 1)), torch.float32), pin((n, max(wa.n_term, 1)), torch.uint8)
    act_h = pin((n, max(wa.n_act, 1)), torch.float32)[:, :wa.n_act]
    for k in range(5):
        act = rng.uniform(lo, hi, (n, wa.n_act)).astype(np.float32)
        np.copyto(act_h, act)
        wa.step_host(act_h if wa.n_act else None, obs_h if wa.n_obs else None, rew_h if wa.n_rew else None, term_h if wa.n_term else None)
        wb.action.copy_(torch.from_numpy(act))
        wb.step()
        torch.cuda.synchronize()
        assert torch.equal(wa.state, wb.state)
        if wa.n_obs:
            assert np.array_equal(obs_h[:, :wa.n_obs], wb.obs.cpu().numpy())
        if wa.n_rew:
            assert np.array_equal(rew_h[:, :wa.n_rew], wb.reward.cpu().numpy())
        if wa.n_term:
            assert np.array_equal(term_h[:, :wa.n_term], wb.term.cpu().numpy())
    a.close()
    b.close()


# ---------------------------------------------------------------- contact capacity ----------------------------------------
def _capacity(factory):
    """max_contacts = 4 on the maze: the R2D2's four wheels alone fill the list, every further contact is counted as dropped and
    the list keeps the deepest ones - on both arms."""
    import yaml
    node = yaml.load(open(os.path.join(ROOT, CONFIGS['r2d2_maze'][0])), Loader=yaml.FullLoader)
    node['max_contacts'] = 4
    env = DIYGym(Configuration.from_dict('r2d2_maze', node), num_envs=2, device=0, world_factory=factory)
    o = OracleWorld(env.scene, env_id=0)
    o.env_reset()
    a = np.array([8.0, 8.0, 10.0, 10.0])
    for _ in range(60):
        env.world.action[:] = torch.from_numpy(a.astype(np.float32))
        env.world.step()
        o.env_step(a)
    assert o.contacts_dropped() > 0 and env.world.contacts_dropped() > 0
    assert len(o.contacts()) <= 4
    assert np.allclose(env.world.state[0, :40].cpu().numpy(), o.state[:40], atol=5e-3)
    env.close()


def test_contact_capacity_counts_and_evicts_cpu():
    _capacity(_factory(8))


@pytest.mark.gpu
def test_contact_capacity_counts_and_evicts_gpu():
    _capacity(None)


# ---------------------------------------------------------------- auto_reset --------------------------------------------
def _auto_reset(factory):
    """with auto_reset the reward returned for a finished environment is the terminal step's, not the post-reset one."""
    import yaml
    node = yaml.load(open(os.path.join(ROOT, CONFIGS['ur_high_5'][0])), Loader=yaml.FullLoader)
    node['max_episode_steps'] = 3
    a = DIYGym(Configuration.from_dict('ur_high_5', node), num_envs=4, device=0, world_factory=factory, auto_reset=True)
    b = DIYGym(Configuration.from_dict('ur_high_5', node), num_envs=4, device=0, world_factory=factory, auto_reset=False)
    g = torch.Generator(device=a.world.action.device).manual_seed(0)
    for k in range(3):
        act = a.sample_action(g)
        oa, ra, ta, info = a.step(act)
        ob, rb, tb, _ = b.step(act)
    assert bool(torch.as_tensor(ta).all()) and bool(torch.as_tensor(tb).all())           # the episode timer fired everywhere
    assert torch.equal(ra['ur_high_5']['reach_goal'], rb['ur_high_5']['reach_goal'])       # the terminal step's reward
    term_obs = info['terminal_observation']['ur5_l']['joint_state']['position']
    assert torch.equal(term_obs, ob['ur5_l']['joint_state']['position'])
    assert not torch.equal(oa['ur5_l']['joint_state']['position'], term_obs)               # the returned observation is the reset one
    assert float(a.step_counter.max()) == 0.0
    a.close()
    b.close()


def test_auto_reset_returns_terminal_reward_cpu():
    _auto_reset(_factory())


@pytest.mark.gpu
def test_auto_reset_returns_terminal_reward_gpu():
    _auto_reset(None)


# ---------------------------------------------------------------- config 5: per-environment parameters -----------------------
def _config5(factory):
    """BASELINE.json config 5: link masses x log-U[0.25, 4], lateral friction U[0.5, 1.25], joint damping x log-U[0.2, 20]
    of the nominal 0.1, base pose +-5 cm / +-0.1 rad - per environment, from the (seed, global environment id, reset count) stream,
    identical on the oracle."""
    n = 32
    env = _env('ur_high_5_randomised', n, factory, seed=21)
    w, sc, h = env.world, env.scene, env.scene.hdr
    p = w.param.cpu().numpy()
    nb, nl, ns, nd = sc['nb'], sc['nl'], sc['ns'], sc['nd']
    fr = p[:, h['P_FRICTION']:h['P_FRICTION'] + ns]
    assert fr.min() >= 0.5 and fr.max() <= 1.25 and fr[:, 0].std() > 0.1          # drawn per environment
    shape_body = sc.sec['SHAPE_I'][:, 0]
    for b in range(nb):
        cols = np.nonzero(shape_body == b)[0]
        assert np.allclose(fr[:, cols], fr[:, cols[:1]])                            # one draw per body
    damp = p[:, h['P_JDAMP']:h['P_JDAMP'] + nd]
    assert damp.min() >= 0.1 * 0.2 * 0.999 and damp.max() <= 0.1 * 20 * 1.001 and damp.std() > 0.1
    mass = p[:, h['P_MASS']:h['P_MASS'] + nb + nl]
    nominal = sc.sec['PARAM_DEFAULT'][h['P_MASS']:h['P_MASS'] + nb + nl]
    ratio = mass[:, nominal > 0] / nominal[nominal > 0]
    assert ratio.min() >= 0.25 * 0.999 and ratio.max() <= 4.0 * 1.001
    pos = w.state[:, h['S_BPOS']:h['S_BPOS'] + 3 * nb].cpu().numpy().reshape(n, nb, 3)
    init = sc.sec['PARAM_DEFAULT'][h['P_INITPOSE']:h['P_INITPOSE'] + 7 * nb].reshape(nb, 7)[:, :3]
    assert np.abs(pos - init).max() <= 0.05 + 1e-6 and np.abs(pos - init)[..., :2].std() > 0.01
    for i in (0, 7, 31):
        o = OracleWorld(sc, seed=21, env_id=i)
        o.env_reset()
        assert np.allclose(p[i], o.param, rtol=1e-5, atol=1e-7)
    env.close()


def test_config5_randomised_parameters_cpu():
    _config5(_factory())


@pytest.mark.gpu
def test_config5_randomised_parameters_gpu():
    _config5(None)


# ---------------------------------------------------------------- visual_randomizer ------------------------------------------
def _visual_randomizer(factory):
    """misc/visual_randomizer.py as per-environment colour randomisation: every reset gives the model's visual shapes
    new colours, different per environment, the same as the oracle's - and the camera shows them."""
    import yaml
    node = yaml.load(open(os.path.join(ROOT, CONFIGS['basic_env'][0])), Loader=yaml.FullLoader)
    node['red_marble']['new_look'] = {'addon': 'visual_randomizer'}
    n = 3
    env = DIYGym(Configuration.from_dict('basic_env', node), num_envs=n, device=0, world_factory=factory, seed=5)
    w, sc, h = env.world, env.scene, env.scene.hdr
    names = [b.name for b in sc.bodies]
    red = names.index('red_marble')
    vis = [v for v in range(sc['nv']) if sc.sec['VIS_I'][v][3] == red]
    col = lambda: w.param[:, h['P_COLOR']:h['P_COLOR'] + 3 * sc['nv']].cpu().numpy().reshape(n, -1, 3)[:, vis]
    c0 = col()
    assert c0.min() >= 0 and c0.max() <= 1 and np.abs(c0[0] - c0[1]).max() > 1e-3            # per environment
    others = [v for v in range(sc['nv']) if v not in vis]
    allc = w.param[:, h['P_COLOR']:h['P_COLOR'] + 3 * sc['nv']].cpu().numpy().reshape(n, -1, 3)
    assert np.allclose(allc[:, others], sc.sec['VIS_F'][others, 11:14][None])                   # the other models keep their colours
    for i in range(n):
        o = OracleWorld(sc, seed=5, env_id=i)
        o.env_reset()
        assert np.allclose(o.param, w.param[i].cpu().numpy(), rtol=1e-6, atol=1e-7)
    env.reset()
    assert np.abs(col() - c0).max() > 1e-3                                                      # a new look on every reset
    # the camera sees the new colours: the marble's pixels carry its colour (times the shading factor)
    rgb = env.observe()['basic_env']['camera']['rgb'][0].cpu().numpy()
    o = OracleWorld(sc, seed=5, env_id=0)
    o.state[:] = w.state[0].cpu().numpy()
    o.param[:] = w.param[0].cpu().numpy()
    o.refresh()
    assert np.abs(rgb - o.render(0)[0]).max() < 2e-3
    env.close()


def test_visual_randomizer_cpu():
    _visual_randomizer(_factory())


@pytest.mark.gpu
def test_visual_randomizer_gpu():
    _visual_randomizer(None)

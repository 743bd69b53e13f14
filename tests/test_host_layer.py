"""Host logic on CPU: the DIYGym facade, Model, add-on registry, spaces and (un)flatten, run against the test-only
CPU build of the kernel code (tests/emul).  Mirrors what the reference's own tests check
(diy_gym/tests/test_config.py, test_utils.py, test_environment.py) plus the structure rules of SURVEY App. F."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from diy_gym_b200 import Configuration, DIYGym
from diy_gym_b200.addons.addon import Addon, AddonFactory
from diy_gym_b200 import spaces
from diy_gym_b200.utils import flatten, unflatten
from tests.emul.world import factory

EX = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'examples')


def make(name, n=2, **kw):
    return DIYGym(os.path.join(EX, name, name + '.yaml'), num_envs=n, world_factory=factory(), **kw)


# ---- reference tests/test_config.py ---------------------------------------------------------------------------
def test_config_get_set_find_all():
    config = Configuration.from_file(os.path.join(EX, 'basic_env', 'basic_env.yaml'))
    assert config.get('im_a_config') is True
    assert config.get('im_not_a_config', False) is False
    config.set('im_a_new_config', 10)
    assert config.get('im_a_new_config') == 10
    assert len(list(config.find_all('model'))) == 4
    with pytest.raises(KeyError, match="Couldn't find config and no default provided"):
        config.get('nope')
    assert config.name == 'basic_env'


def test_missing_urdf_raises_value_error():
    cfg = Configuration.from_dict('x', {'thing': {'model': 'does/not/exist.urdf'}})
    with pytest.raises(ValueError, match='Could not find URDF: does/not/exist.urdf'):
        DIYGym(cfg, num_envs=1, world_factory=factory())


# ---- reference tests/test_environment.py ----------------------------------------------------------------------
def test_basic_env_models_and_spaces():
    env = make('basic_env')
    for m in ('plane', 'red_marble', 'green_marble', 'blue_marble'):
        assert m in env.models
    assert 'force' in env.action_space['blue_marble'].spaces
    assert 'camera' in env.observation_space['basic_env'].spaces
    assert 'pose' in env.observation_space['green_marble'].spaces
    assert env.observation_space['basic_env']['camera']['rgb'].shape == (50, 50, 3)
    assert list(env.receptors.keys()) == ['basic_env', 'blue_marble', 'green_marble', 'plane', 'red_marble']


def test_basic_env_episode_marble_moves_and_reset_restores():
    env = make('basic_env', n=2)
    obs = env.reset()
    initial = obs['green_marble']['pose']['position'].clone()
    push = torch.tensor([[0., -100., 0.], [0., -100., 0.]])
    for _ in range(500):
        obs, _, _, _ = env.step({'blue_marble': {'force': push}})
    final = obs['green_marble']['pose']['position']
    assert abs(float(initial[0].norm()) - float(final[0].norm())) > 0.5     # the green marble has moved
    obs = env.reset()
    assert abs(float(initial[0].norm()) - float(obs['green_marble']['pose']['position'][0].norm())) < 0.05
    assert obs['basic_env']['camera']['rgb'].shape == (2, 50, 50, 3)


def test_camera_segmentation_mask_reports_body_ids():
    """sensors/camera.py:54-56,89-90: `use_segmentation_mask` adds a per-pixel unique-id image (-1 = background)."""
    import copy
    import yaml
    cfg = yaml.safe_load(open(os.path.join(EX, 'basic_env', 'basic_env.yaml')))
    cfg['camera']['use_segmentation_mask'] = True
    path = os.path.join(os.path.dirname(__file__), '_tmp_seg.yaml')
    with open(path, 'w') as f:
        yaml.safe_dump(cfg, f)
    try:
        env = DIYGym(path, num_envs=1, world_factory=factory())
        assert env.observation_space['_tmp_seg']['camera']['segmentation_mask'].shape == (50, 50)
        obs = env.reset()
        mask = obs['_tmp_seg']['camera']['segmentation_mask'][0].numpy()
        ids = {int(v) for v in np.unique(mask)}
        marbles = {env.models[m].body.index for m in ('red_marble', 'green_marble', 'blue_marble')}
        assert marbles <= ids and env.models['plane'].body.index in ids      # the overhead camera sees the plane and the three marbles
        assert ids <= marbles | {env.models['plane'].body.index, -1}
    finally:
        os.remove(path)


def test_camera_rgb_uint8_key_returns_the_renderers_bytes():
    """Extension key `rgb_uint8`: the colour observation as bytes round(255 c) - what the reference divides by 255 (camera.py:76-78)."""
    import yaml
    cfg = yaml.safe_load(open(os.path.join(EX, 'basic_env', 'basic_env.yaml')))
    path = os.path.join(os.path.dirname(__file__), '_tmp_u8.yaml')
    try:
        imgs = {}
        for u8 in (False, True):
            cfg['camera']['rgb_uint8'] = u8
            with open(path, 'w') as f:
                yaml.safe_dump(cfg, f)
            env = DIYGym(path, num_envs=1, world_factory=factory())
            sp = env.observation_space['_tmp_u8']['camera']['rgb']
            assert sp.dtype == (np.uint8 if u8 else np.float32) and sp.shape == (50, 50, 3)
            imgs[u8] = env.reset()['_tmp_u8']['camera']['rgb'][0].numpy()
        assert imgs[True].dtype == np.uint8
        assert np.array_equal(imgs[True], np.rint(np.clip(imgs[False], 0, 1) * 255).astype(np.uint8))
    finally:
        if os.path.exists(path):
            os.remove(path)


# ---- reference tests/test_utils.py ----------------------------------------------------------------------------
def test_flatten_unflatten_round_trip():
    env = make('basic_env')
    action = env.sample_action()
    flat = flatten(action)
    assert flat.shape == (2, 6)
    back = unflatten(flat, env.action_space)
    assert torch.equal(action['red_marble']['force'], back['red_marble']['force'])
    assert torch.equal(action['blue_marble']['force'], back['blue_marble']['force'])
    # per-environment numpy form, as the reference uses it
    a1 = env.action_space.sample()
    b1 = unflatten(torch.as_tensor(flatten(a1, batched=False)), env.action_space, batched=False)
    assert np.allclose(a1['red_marble']['force'], b1['red_marble']['force'])


# ---- structure of every example -----------------------------------------------------------
def test_ur_high_5_trees():
    env = make('ur_high_5', n=3)
    assert list(env.action_space.spaces) == ['ur5_l', 'ur5_r']
    assert list(env.action_space['ur5_l']['controller'].spaces) == ['linear', 'rotation']
    assert list(env.observation_space.spaces) == ['ur5_l', 'ur5_r', 'ur_high_5']
    js = env.observation_space['ur5_l']['joint_state']
    assert js['position'].shape == (6, ) and np.allclose(js['velocity'].high, [3.15, 3.15, 3.15, 3.2, 3.2, 3.2])
    assert np.allclose(js['position'].high, [2 * np.pi] * 2 + [np.pi] + [2 * np.pi] * 3, atol=1e-3)
    obs, rew, term, info = env.step(env.sample_action())
    assert list(obs.keys()) == ['ur5_l', 'ur5_r', 'ur_high_5'] and obs['ur5_l']['joint_state']['position'].shape == (3, 6)
    assert obs['ur_high_5']['distance_to_target']['position'].shape == (3, 3)
    assert list(rew.keys()) == ['ur_high_5'] and rew['ur_high_5']['reach_goal'].shape == (3, )
    assert term.dtype == torch.bool and term.shape == (3, )      # terminal_if_any collapses to one bool per environment
    # reward is -|ee_r - ee_l| in the world frame, the sensor reports the same vector
    d = obs['ur_high_5']['distance_to_target']['position'].norm(dim=1)
    assert torch.allclose(-d, rew['ur_high_5']['reach_goal'], atol=1e-6)
    assert info == {}


def test_from_the_readme_trees():
    env = make('from_the_readme', n=1)
    assert list(env.receptors) == ['from_the_readme', 'plane', 'r2d2', 'robot', 'table']
    assert list(env.action_space['robot']['controller'].spaces) == ['linear']
    obs, rew, term, _ = env.step(env.sample_action())
    assert obs['r2d2']['arm_camera']['rgb'].shape == (1, 200, 200, 3) and obs['r2d2']['arm_camera']['depth'].shape == (1, 200, 200)
    assert float(obs['r2d2']['arm_camera']['depth'].max()) < 0      # eye-space z is negative, as the reference yields
    assert list(rew.keys()) == ['from_the_readme', 'robot'] and list(rew['robot']) == ['lazy_robot']
    assert list(term['from_the_readme'].keys()) == ['grab_r2d2', 'episode_timer']
    assert float(rew['robot']['lazy_robot'][0]) <= 0


def test_r2d2_maze_wheel_order_follows_joint_index():
    env = make('r2d2_maze', n=1)
    assert env.scene['nb'] == 121                                    # plane + R2D2 + 119 walls
    ctrl = env.models['r2d2'].addons['wheel_driver']
    names = [env.models['r2d2'].body.joint_names()[i] for i in ctrl.joint_ids]
    assert names == ['right_front_wheel_joint', 'right_back_wheel_joint', 'left_front_wheel_joint', 'left_back_wheel_joint']
    assert env.action_space['r2d2']['wheel_driver'].shape == (4, )
    assert len(env.observation_space.spaces) == 0
    x0 = env.models['r2d2'].base_pose()[0].clone()
    for _ in range(60):
        obs, rew, term, _ = env.step({'r2d2': {'wheel_driver': torch.full((1, 4), 0.5) * 20}})
    assert obs == {} and rew == {} and term == {}
    assert float((env.models['r2d2'].base_pose()[0] - x0)[0, :2].norm()) > 0.05   # it drives


def test_episode_timer_and_masked_reset():
    env = make('from_the_readme', n=2)
    env.world.state[:, env.scene.hdr['S_STEP']] = 498
    _, _, term, _ = env.step(env.sample_action())
    assert not bool(term['from_the_readme']['episode_timer'].any())
    _, _, term, _ = env.step(env.sample_action())
    assert bool(term['from_the_readme']['episode_timer'].all())
    env.reset(torch.tensor([True, False]))
    assert env.step_counter.tolist() == [0.0, 500.0]


def test_options_sum_rewards_flatten_hide():
    cfg = Configuration.from_file(os.path.join(EX, 'ur_high_5', 'ur_high_5.yaml'))
    cfg.set('sum_rewards', True)
    cfg.set('flatten_observations', True)
    cfg.set('flatten_actions', True)
    cfg.node['ur5_r']['joint_state']['hide'] = True
    env = DIYGym(cfg, num_envs=2, world_factory=factory())
    assert isinstance(env.action_space, spaces.Box) and env.action_space.shape == (12, )
    assert env.observation_space.shape == (15, )                    # hidden add-on leaves the SPACE ...
    obs, rew, term, _ = env.step(torch.zeros(2, 12))
    assert obs.shape == (2, 27)                                     # ... but still observes
    assert rew.shape == (2, ) and term.shape == (2, )


def test_actions_absent_from_the_dict_are_not_applied():
    env = make('basic_env', n=1)
    env.reset()
    push = torch.tensor([[0., -100., 0.]])
    env.step({'blue_marble': {'force': push}})
    v1 = env.models['blue_marble'].base_velocity()[0][0, 1].item()
    for _ in range(5):
        env.step({})                                                # stale force in the buffer must not act again
    v2 = env.models['blue_marble'].base_velocity()[0][0, 1].item()
    assert v1 < -0.01 and abs(v2) <= abs(v1)


def test_user_addons_register_and_run():
    spec = importlib.util.spec_from_file_location('drone_pilot_example', os.path.join(EX, 'drone_pilot', 'drone_pilot.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert AddonFactory.get().addons['propellor'] is mod.Propellor
    env = make('drone_pilot', n=3)
    assert list(env.models['drone'].addons) == ['fell_over', 'motor1', 'motor2', 'motor3', 'motor4', 'pose', 'respawn']
    assert list(env.observation_space['drone']['pose'].spaces) == ['position', 'rotation', 'velocity', 'angular_velocity']
    obs = env.reset()
    assert list(obs['drone']['pose'].keys()) == ['position', 'velocity', 'rotation', 'angular_velocity']   # insertion order
    tp = obs['target']['pose']['position']
    assert float((tp[0] - tp[1]).norm()) > 1e-3                      # per-environment respawn draws differ
    assert bool(((tp - tp.new_tensor([0, 0, 5.0])).abs() <= tp.new_tensor([5.0, 5.0, 1.0]) + 1e-5).all())
    full = {'drone': {'motor%d' % (i + 1): torch.full((3, 1), 1.5) for i in range(4)}}   # the reference does not clip either
    for _ in range(150):
        obs, rew, term, _ = env.step(full)
    assert float(obs['drone']['motor1'].min()) > 1.49                # spool-up state lives in the add-on
    assert float(obs['drone']['pose']['position'][:, 2].min()) > 1.0  # 4 x 30 N lifts the 8 kg drone (base 4 + 4 motor links of 1)
    assert term.shape == (3, ) and rew['drone_pilot']['reach_goal'].shape == (3, )


def test_custom_addon_can_be_defined_inline():
    class Constant(Addon):
        def __init__(self, parent, config):
            super().__init__(parent, config)
            self.observation_space = spaces.Box(0, 1, shape=(2, ))

        def observe(self):
            return torch.full((self.parent.env.num_envs, 2), 0.25)

        def reward(self):
            return torch.ones(self.parent.env.num_envs)
    AddonFactory.register_addon('constant', Constant)
    cfg = Configuration.from_dict('tiny', {'sum_rewards': True, 'ball': {'model': 'sphere2.urdf', 'xyz': [0, 0, 1], 'c': {'addon': 'constant'}}})
    env = DIYGym(cfg, num_envs=2, world_factory=factory())
    obs, rew, term, _ = env.step({})
    assert torch.equal(obs['ball']['c'], torch.full((2, 2), 0.25)) and torch.equal(rew, torch.ones(2))


# ---- admittance_controller, force_torque_sensor ----------------------------------------------
def test_admittance_controller_and_force_torque_sensor_trees():
    """Same spaces as admittance_controller.py:26-28 / force_torque_sensor.py:16-19; the controller holds the arm against
    gravity (zero wrench -> the tool barely moves) and a commanded force pushes the tool along it."""
    env = make('ur_admittance', n=2)
    ctl = env.action_space['ur5']['controller']
    assert list(ctl.spaces) == ['force', 'torque']
    assert np.allclose(ctl['force'].high, 5) and np.allclose(ctl['torque'].high, 1) and ctl['force'].shape == (3, )
    ft = env.observation_space['ur5']['wrist_wrench']
    assert list(ft.spaces) == ['force', 'torque'] and ft['force'].shape == (3, )
    zero = {'ur5': {'controller': {'force': torch.zeros(2, 3), 'torque': torch.zeros(2, 3)}}}
    obs0 = env.reset()
    p0 = obs0['ur5']['tool_state']['position'].clone()
    for _ in range(20):
        obs, rew, term, _ = env.step(zero)
    # (the hot-start step of reset() runs without the controller, as in the reference, so the arm starts with a small sag velocity)
    assert (obs['ur5']['tool_state']['position'] - p0).abs().max() < 1e-2
    w = obs['ur5']['wrist_wrench']
    assert w['force'].shape == (2, 3) and 1.0 < w['force'][0].norm() < 50.0           # carries the wrist-3 link + tool
    push = {'ur5': {'controller': {'force': torch.tensor([[5.0, 0, 0], [0, 0, 0]]), 'torque': torch.zeros(2, 3)}}}
    for _ in range(40):
        obs, rew, term, _ = env.step(push)
    moved = obs['ur5']['tool_state']['position'] - p0
    assert moved[0, 0] > 5e-3 and moved[0, 0] > 5 * abs(moved[1, 0]) and moved[1].abs().max() < 2e-2


def test_admittance_controller_rejects_partial_chains_like_pybullet():
    import yaml
    cfg = {'ur5': {'model': 'ur5/ur5_robot.urdf', 'controller': {'addon': 'admittance_controller', 'end_effector': 'elbow_joint'}}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_tmp_admittance.yaml')
    with open(path, 'w') as f:
        yaml.safe_dump(cfg, f)
    try:
        with pytest.raises(ValueError):
            make_path(path)
    finally:
        os.remove(path)


def make_path(path, n=1):
    return DIYGym(path, num_envs=n, world_factory=factory())


# ---- nested models welded to a parent frame (model.py:69-77) -----------------------------------------
def test_nested_model_is_welded_to_the_parent_frame():
    env = DIYGym(os.path.join(EX, 'ur_gripper', 'ur_gripper.yaml'), num_envs=2, world_factory=factory(team=8))
    ur5 = env.models['ur5']
    payload = ur5.models['payload']
    assert env.scene['ncons'] == 1 and payload.uid == 1
    # like the reference, only top-level models are receptors: the nested model's add-ons are built but never walked
    assert list(env.observation_space.spaces) == ['ur5'] and list(env.observation_space['ur5'].spaces) == ['joint_state']
    env.reset()
    act = {'ur5': {'controller': {'linear': torch.tensor([[0.01, 0.0, 0.0], [0.0, 0.0, 0.0]])}}}
    ee = ur5.get_frame_id('ee_fixed_joint')
    for _ in range(80):
        env.step(act)
    pos, quat, _, _ = ur5.link_state(ee)
    from diy_gym_b200 import torch_math as tm
    target = pos + tm.quat_rotate(quat, pos.new_tensor([0.0, 0.0, 0.1]).expand_as(pos))
    gap = (payload.base_pose()[0] - target).norm(dim=1)
    assert gap.max() < 0.02                      # held by the 6 constraint rows (soft: erp 0.2 per sub-step, under gravity)
    assert pos[0, 0] - pos[1, 0] > 0.02          # environment 0 was driven along +x, its payload came along
    assert (payload.base_pose()[0][0, 0] - payload.base_pose()[0][1, 0]) > 0.02


def test_drone_pilot_lowered_user_addons_equal_the_torch_forms():
    """examples/drone_pilot: `Propellor` / `FellOver` as ops of the fused step (the default) against the same add-ons as eager
    PyTorch (`propellor_torch`, `fell_over_torch` - the general path for arbitrary user code): same observations, rewards and
    terminals over a rollout.  (CPU build of the kernel source; the GPU leg is tests/test_gpu_parity.py.)"""
    import yaml
    from bench import CONFIGS, register_example_addons
    from tests.emul.world import factory
    register_example_addons()
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
    node = yaml.load(open(os.path.join(root, CONFIGS['drone_pilot'][0])), Loader=yaml.FullLoader)
    envs = []
    for torch_form in (False, True):
        nd = yaml.load(yaml.dump(node), Loader=yaml.FullLoader)
        if torch_form:
            for k, v in nd['drone'].items():
                if isinstance(v, dict) and v.get('addon') in ('propellor', 'fell_over'):
                    v['addon'] += '_torch'
        envs.append(DIYGym(Configuration.from_dict('drone_pilot', nd), num_envs=3, world_factory=factory(4), seed=11))
    a, b = envs
    assert a.scene['nop'] > b.scene['nop']                      # four FILTERED_WRENCH ops and one TILT_TERMINAL more
    g = torch.Generator().manual_seed(5)
    for k in range(25):
        act = {'drone': {'motor%d' % (i + 1): torch.rand((3, 1), generator=g) * (1.5 if k > 10 else 0.3) for i in range(4)}}
        oa, ra, ta, _ = a.step(act)
        ob, rb, tb, _ = b.step(act)
        for m in ('motor1', 'motor4'):
            assert torch.allclose(oa['drone'][m], ob['drone'][m], atol=1e-6)
        assert torch.allclose(oa['drone']['pose']['position'], ob['drone']['pose']['position'], atol=1e-5)
        assert torch.allclose(oa['drone']['pose']['rotation'], ob['drone']['pose']['rotation'], atol=1e-5)
        assert torch.equal(ta, tb)
    assert float(oa['drone']['motor1'].min()) > 0.3                  # the rotors spooled up

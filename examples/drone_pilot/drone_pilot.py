"""drone_pilot on the batched backend: N quadrotors, each chasing its own random target.

Shows the user add-on surface: `Propellor` (stateful, link-frame thrust + torque) and `FellOver` (terminal) are registered
exactly like in the reference (`AddonFactory.register_addon`) and used from the unchanged YAML.  They come in two forms: lowered
to ops of the fused CUDA step (subclasses of the `filtered_link_wrench` / `tilt_terminal` building blocks - the default) and as
plain batched torch code on the views the parent Model exposes (`PropellorTorch`, `FellOverTorch`: the general path for
arbitrary user code).
Semantics follow the reference's examples/drone_pilot/drone_pilot.py:10-59.

    python examples/drone_pilot/drone_pilot.py [num_envs] [steps]
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from diy_gym_b200 import DIYGym, spaces  # noqa: E402
from diy_gym_b200 import torch_math as tm  # noqa: E402
from diy_gym_b200.addons.addon import Addon, AddonFactory  # noqa: E402
from diy_gym_b200.addons.builtin import FilteredLinkWrench, TiltTerminal  # noqa: E402


class Propellor(FilteredLinkWrench):
    """First-order rotor spool-up, thrust along the motor link's z axis and a reaction torque about it - the reference's
    user add-on (examples/drone_pilot/drone_pilot.py:10-40), LOWERED to one op of the fused step: a user add-on may emit scene
    ops from compile() instead of running Python every step.  Same config keys as the reference's class."""
    def __init__(self, parent, config):
        super().__init__(parent, config)
        self.max_thrust = config.get('max_thrust', 20.0)
        self.max_torque = config.get('max_torque', 0.1) * (1.0 if config.get('rotor_direction') == 'CCW' else -1.0)
        self.spool_up_rate = 0.1
        self.force, self.torque, self.rate = [0.0, 0.0, float(self.max_thrust)], [0.0, 0.0, float(self.max_torque)], self.spool_up_rate


class FellOver(TiltTerminal):
    """Terminal when the base is tilted by more than 10 degrees (drone_pilot.py:43-55 of the reference), as an op."""


class PropellorTorch(Addon):
    """The same add-on as plain batched PyTorch on the views the parent Model exposes - the general path for arbitrary user
    code (a few small torch kernels per add-on per step).  Registered as `propellor_torch`; DG_DRONE_TORCH=1 makes the example
    YAML use it (A/B against the lowered form: tests/test_host_layer.py)."""
    def __init__(self, parent, config):
        super().__init__(parent, config)
        self.frame_id = parent.get_frame_id(config.get('frame'))
        self.max_thrust = config.get('max_thrust', 20.0)
        self.max_torque = config.get('max_torque', 0.1) * (1.0 if config.get('rotor_direction') == 'CCW' else -1.0)
        self.spool_up_rate = 0.1
        self.rotor_speed = None   # [num_envs, 1], created once the device world exists
        self.observation_space = spaces.Box(0.0, 1.0, shape=(1, ), dtype='float32')
        self.action_space = spaces.Box(0.0, 1.0, shape=(1, ), dtype='float32')

    def bind(self, env):
        self.rotor_speed = torch.zeros((env.num_envs, 1), device=env.world.state.device)

    def update(self, action):
        a = torch.as_tensor(action, device=self.rotor_speed.device, dtype=torch.float32).reshape(-1, 1)
        self.rotor_speed += (a - self.rotor_speed) * self.spool_up_rate
        z = torch.zeros_like(self.rotor_speed)
        self.parent.apply_external_force(self.frame_id, torch.cat([z, z, self.max_thrust * self.rotor_speed], 1), frame='link')
        self.parent.apply_external_torque(self.frame_id, torch.cat([z, z, self.max_torque * self.rotor_speed], 1), frame='link')

    def observe(self):
        return self.rotor_speed


class FellOverTorch(Addon):
    """Terminal when the base is tilted by more than 10 degrees, in PyTorch."""
    def is_terminal(self):
        quat = self.parent.base_pose()[1]
        return 2.0 * torch.atan2(quat[:, :3].norm(dim=1), quat[:, 3].abs()) > math.radians(10)


_TORCH = os.environ.get('DG_DRONE_TORCH', '0') == '1'
AddonFactory.register_addon('propellor', PropellorTorch if _TORCH else Propellor)
AddonFactory.register_addon('fell_over', FellOverTorch if _TORCH else FellOver)
AddonFactory.register_addon('propellor_torch', PropellorTorch)
AddonFactory.register_addon('fell_over_torch', FellOverTorch)

# thrust, roll, pitch, yaw torque -> four motor speeds (same mixer as the reference's example)
MIXER = torch.linalg.inv(torch.tensor([[1., 1, 1, 1], [0, -1, 0, 1], [1, 0, -1, 0], [1, -1, 1, -1]]))


def pilot(env, steps=2000):
    """Cascaded P / PID controller flying every drone to its target; returns the fraction that got there."""
    dev = env.world.state.device
    n = env.num_envs
    mixer = MIXER.to(dev)
    obs = env.reset()
    integral = torch.zeros(n, device=dev)
    reached = torch.zeros(n, dtype=torch.bool, device=dev)
    for _ in range(steps):
        d, t = obs['drone']['pose'], obs['target']['pose']
        q = tm.quat_from_euler(d['rotation'])
        err = tm.quat_rotate_inv(q, t['position'] - d['position'])
        roll, pitch = err[:, 1] * 0.0005, -err[:, 0] * 0.0005
        q_ref = tm.quat_from_euler(torch.stack([roll, pitch, torch.zeros_like(roll)], 1))
        integral += err[:, 2]
        thrust = err[:, 2] * 0.75 + integral * 0.00075 - d['velocity'][:, 2] * 0.5
        torque = -tm.quat_mul(q_ref, q)[:, :3] * 0.01
        torque = torch.where((d['position'][:, 2] < 2.0).unsqueeze(1), torch.zeros_like(torque), torque)
        speeds = (mixer @ torch.stack([thrust, torque[:, 0], torque[:, 1], torque[:, 2]], 0)).clamp(0, 1)
        action = {'drone': {'motor%d' % (i + 1): speeds[i].unsqueeze(1) for i in range(4)}}
        obs, reward, terminal, _ = env.step(action)
        reached |= obs['drone']['pose']['position'].sub(obs['target']['pose']['position']).norm(dim=1) < 0.1
        if bool(terminal.any()):
            integral[terminal] = 0
            obs = env.reset(terminal)
    return float(reached.float().mean())


if __name__ == '__main__':
    num_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    env = DIYGym(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'drone_pilot.yaml'), num_envs=num_envs)
    print('fraction of drones that reached their target: %.2f' % pilot(env, steps))

"""from_the_readme on the batched backend: random actions on N environments, masked reset on terminal.

    python examples/from_the_readme/from_the_readme.py [num_envs] [steps]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from diy_gym_b200 import DIYGym  # noqa: E402
from diy_gym_b200.utils import walk_dict  # noqa: E402

if __name__ == '__main__':
    num_envs = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    env = DIYGym(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'from_the_readme.yaml'), num_envs=num_envs)
    obs = env.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    episodes = 0
    for _ in range(steps):
        obs, reward, terminal, _ = env.step(env.sample_action())
        done = terminal if isinstance(terminal, torch.Tensor) else walk_dict(terminal, any) if len(terminal) else None
        if isinstance(done, torch.Tensor) and bool(done.any()):
            episodes += int(done.sum())
            env.reset(done)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print('%s: %d envs x %d steps in %.2f s = %.0f env-steps/s, %d episodes finished' % ('from_the_readme', num_envs, steps, dt, num_envs * steps / dt, episodes))

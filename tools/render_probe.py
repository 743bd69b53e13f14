"""Times the camera kernel alone: python tools/render_probe.py [config] [n_envs]"""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from bench import CONFIGS, ROOT, register_example_addons  # noqa: E402
from diy_gym_b200 import DIYGym  # noqa: E402

register_example_addons()
name = sys.argv[1] if len(sys.argv) > 1 else 'from_the_readme'
n = int(sys.argv[2]) if len(sys.argv) > 2 else CONFIGS[name][1]
env = DIYGym(os.path.join(ROOT, CONFIGS[name][0]), num_envs=n)
w = env.world
for _ in range(3):
    w.step()
for _ in range(3):
    w.render(0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    w.render(0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
wd, hg = w.cams[0]
print('%s n %d camera %dx%d: %.3f ms per render, %.1f Gpixel/s, output %.0f GB/s' % (name, n, wd, hg, ms, n * wd * hg / ms / 1e6, n * wd * hg * 16 / ms / 1e6))

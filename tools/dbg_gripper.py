import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from diy_gym_b200 import DIYGym
from diy_gym_b200.backend import World
from oracle.oracle import OracleWorld
np.set_printoptions(linewidth=200, precision=6, suppress=True)
env = DIYGym('examples/ur_gripper/ur_gripper.yaml', num_envs=1, compile_only=True)
sc = env.scene
h = sc.hdr
g = np.load('tests/golden/ur_gripper.npz')
o = OracleWorld(sc, seed=int(g['seed']), env_id=0); obs, _, _ = o.env_reset()
print('golden obs0', g['obs_0'][0])
print('oracle obs ', obs)
w = World(sc, 1, seed=int(g['seed']), env_id_offset=0)
w.reset(); torch.cuda.synchronize()
print('gpu obs    ', w.obs.cpu().numpy()[0])
for k, a in enumerate(g['actions'][:3]):
    w.action.copy_(torch.from_numpy(a[None].astype(np.float32))); w.step(); torch.cuda.synchronize()
    oo = o.env_step(a)[0]
    print(k, 'oracle', oo[6:], '\n  gpu   ', w.obs.cpu().numpy()[0][6:], '\n  golden', g['obs_0'][k + 1][6:])

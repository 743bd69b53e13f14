"""Hand-assembled scenes (SceneBuilder + add-on ops) used by early kernel tests before/without the YAML layer."""
import numpy as np

from diy_gym_b200.assets import resolve_model
from diy_gym_b200.compiler.mathutil import quat_from_euler
from diy_gym_b200.compiler.scene import SceneBuilder


def ik_op(sb, body, ee_joint_name, use_orn, rest):
    """Same bookkeeping as diy_gym/addons/controllers/ik_controller.py:22-45."""
    ee = body.joint_index(ee_joint_name)
    joints = [i for i in body.movable_joints() if i <= ee]
    infos = [body.joint_info(i) for i in joints]
    n, ndb = len(joints), body.n_dofs
    lower = [i['lower'] for i in infos]
    upper = [i['upper'] for i in infos]
    rng = [u - l for l, u in zip(lower, upper)]
    rest = list(rest)[:n] + [0.0] * (n - len(rest))
    nullspace = int(n == ndb)
    pad = lambda v: list(v) + [0.0] * (ndb - len(v))
    fargs = [0.015, 1.0] + [i['max_force'] for i in infos] + pad(lower) + pad(upper) + pad(rng) + pad(rest)
    iargs = [body.index, body.link_start + ee, n, int(use_orn), nullspace] + [body.global_dof(i) for i in joints]
    sb.add_op('IK_CTRL', iargs, fargs, n_act=6 if use_orn else 3)
    sb.add_op('JOINT_RESET', [n] + [body.global_dof(i) for i in joints], rest)


def ur_high_5(max_contacts=16):
    """examples/ur_high_5/ur_high_5.yaml assembled by hand."""
    sb = SceneBuilder(max_contacts=max_contacts)
    desc = resolve_model('ur5/ur5_robot.urdf')
    bl = sb.add_body('ur5_l', desc, xyz=(0, 0.5, 0), quat=quat_from_euler([0, 0, 0]))
    br = sb.add_body('ur5_r', desc, xyz=(0, -0.5, 0), quat=quat_from_euler([0, 0, 0]))
    rest_l = [-0.17, -0.73, -1.93, -0.36, -0.03, -0.06]
    rest_r = [0.17, -2.41, 1.93, -2.78, 0.03, 0.06]
    for b, rest in ((bl, rest_l), (br, rest_r)):
        ik_op(sb, b, 'ee_fixed_joint', True, rest)
    for b in (bl, br):
        dofs = [b.global_dof(i) for i in b.movable_joints()]
        sb.add_op('JOINT_SENSOR', [len(dofs), 1] + dofs, n_obs=2 * len(dofs))
    fl, fr = bl.frame(bl.joint_index('ee_fixed_joint')), br.frame(br.joint_index('ee_fixed_joint'))
    sb.add_op('OBJECT_SENSOR', [fr, fl, 0], n_obs=3)
    sb.add_op('REACH_TARGET', [fl, fr], [1.0, 0.01], n_rew=1, n_term=1)
    return sb.finalize()

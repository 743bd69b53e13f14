"""Shared test helpers: scene construction and an *independent* numpy rigid-body reference
(world-frame Jacobian mass matrix + potential-energy gradient) used to pin the oracle's ABA."""
import numpy as np

from diy_gym_b200.assets import resolve_model
from diy_gym_b200.compiler.mathutil import Transform, quat_rotate, quat_to_mat
from diy_gym_b200.compiler.scene import SceneBuilder


def build_single(model, world=None, **kw):
    sb = SceneBuilder(**(world or {}))
    sb.add_body('m', resolve_model(model), **kw)
    return sb.finalize()


def free_dynamics(w, sc):
    """Switch off motors and every damping term of an OracleWorld."""
    w.s('S_MMAXF', sc['nd'])[:] = 0
    w.p('P_LINDAMP', sc['nb'])[:] = 0
    w.p('P_ANGDAMP', sc['nb'])[:] = 0
    w.p('P_JDAMP', sc['nd'])[:] = 0


def urdf_kinematics(sc, q, body=0):
    """World transforms of every URDF link frame / COM frame as a plain product of the URDF <origin>s."""
    b = sc.bodies[body]
    L, s = b.links, b.scale
    Tl = [None] * len(L)
    Tl[0] = Transform(b.base_pos, b.base_quat) * Transform.from_xyz_rpy(np.array(L[0]['inertial_xyz']) * s,
                                                                       L[0]['inertial_rpy']).inverse()
    axes, orgs, coms, dof = [None] * len(L), [None] * len(L), [None] * len(L), 0
    for k in range(1, len(L)):
        j = L[k]['joint']
        T = Tl[L[k]['parent']] * Transform.from_xyz_rpy(np.array(j['xyz']) * s, j['rpy'])
        if j['type'] in ('revolute', 'continuous'):
            ax = np.array(j['axis']) / np.linalg.norm(j['axis'])
            axes[k], orgs[k] = quat_rotate(T.q, ax), T.p.copy()
            T = T * Transform((0, 0, 0), np.r_[ax * np.sin(q[dof] / 2), np.cos(q[dof] / 2)])
            dof += 1
        elif j['type'] == 'prismatic':
            ax = np.array(j['axis']) / np.linalg.norm(j['axis'])
            axes[k] = quat_rotate(T.q, ax)
            T = T * Transform(ax * q[dof])
            dof += 1
        Tl[k] = T
    for k in range(len(L)):
        coms[k] = Tl[k] * Transform.from_xyz_rpy(np.array(L[k]['inertial_xyz']) * s, L[k]['inertial_rpy'])
    return Tl, coms, axes, orgs


def mass_matrix_gravity_energy(sc, w, q, qd=None, g=(0, 0, -9.81), body=0):
    """M(q), generalized gravity G(q) (M qdd + G = tau at qd = 0) and total energy, fixed-base bodies."""
    b = sc.bodies[body]
    L, nd = b.links, b.n_dofs
    Tl, coms, axes, orgs = urdf_kinematics(sc, q, body)
    mass = w.p('P_MASS', sc['nframes'])
    inert = w.p('P_INERTIA', 3 * sc['nframes']).reshape(-1, 3)
    d, dof_of = 0, [-1] * len(L)
    for k in range(1, len(L)):
        if L[k]['joint']['type'] != 'fixed':
            dof_of[k] = d
            d += 1
    M, G, pot = np.zeros((nd, nd)), np.zeros(nd), 0.0
    for k in range(1, len(L)):
        f = b.frame(k - 1)
        m, Iw = mass[f], None
        R = quat_to_mat(coms[k].q)
        Iw = R @ np.diag(inert[f]) @ R.T
        Jv, Jw = np.zeros((3, nd)), np.zeros((3, nd))
        a = k
        while a > 0:
            if dof_of[a] >= 0:
                if L[a]['joint']['type'] == 'prismatic':
                    Jv[:, dof_of[a]] = axes[a]
                else:
                    Jw[:, dof_of[a]] = axes[a]
                    Jv[:, dof_of[a]] = np.cross(axes[a], coms[k].p - orgs[a])
            a = L[a]['parent']
        M += m * Jv.T @ Jv + Jw.T @ Iw @ Jw
        G += -Jv.T @ (m * np.array(g))
        pot += -m * np.dot(g, coms[k].p)
    energy = pot + (0.5 * qd @ M @ qd if qd is not None else 0.0)
    return M, G, energy

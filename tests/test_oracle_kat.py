"""Tier-0 known-answer tests that pin the CPU oracle without pybullet (SURVEY.md §8c list).

The oracle is "parity unpinned" against the real engine; these tests pin it against closed forms and an
independent numpy rigid-body reference (tests/helpers.py).
"""
import numpy as np
import pytest

from diy_gym_b200.assets import resolve_model
from diy_gym_b200.compiler.mathutil import quat_from_euler, quat_to_mat
from diy_gym_b200.compiler.scene import SceneBuilder
from oracle.oracle import OracleWorld
from tests.helpers import build_single, free_dynamics, mass_matrix_gravity_energy, urdf_kinematics

ARMS = ['ur5/ur5_robot.urdf', 'jaco/j2s7s300_standalone.urdf']


def _valid_q(model, rng, nd):
    q = rng.uniform(-1, 1, nd)
    if 'jaco' in model:
        q[[1, 3, 5]] += 3.0
        q[7:] = rng.uniform(0.2, 1.8, 3)
    return q


@pytest.mark.parametrize('model', ARMS)
def test_fk_matches_urdf_transform_product(model):
    sc = build_single(model, xyz=(0.1, -0.2, 0.3), quat=quat_from_euler([0.3, -0.2, 0.5]))
    w = OracleWorld(sc)
    q = _valid_q(model, np.random.default_rng(0), sc['nd'])
    w.s('S_Q', sc['nd'])[:] = q
    w.refresh()
    Tl, coms, _, _ = urdf_kinematics(sc, q)
    for k in range(1, len(sc.bodies[0].links)):
        fs = w.frame_state(sc.bodies[0].frame(k - 1))
        assert np.allclose(fs['com_pos'], coms[k].p, atol=1e-12)
        assert np.allclose(fs['link_pos'], Tl[k].p, atol=1e-12)
        assert np.allclose(quat_to_mat(fs['link_quat']), quat_to_mat(Tl[k].q), atol=1e-12)


def test_ur5_zero_pose_tool_position():
    """Published UR5 zero-configuration tool flange position (0.81725, 0.19145, -0.005491)."""
    sc = build_single('ur5/ur5_robot.urdf')
    w = OracleWorld(sc)
    w.refresh()
    assert np.allclose(w.frame_state(sc.bodies[0].frame(7))['link_pos'], [0.81725, 0.19145, -0.005491], atol=1e-9)


@pytest.mark.parametrize('model', ARMS)
def test_aba_equals_mass_matrix_solve(model):
    """qdd from the articulated-body algorithm == M(q)^-1 (tau - G(q)) from world-frame Jacobians."""
    dt = 1e-7
    sc = build_single(model, world=dict(timestep=dt))
    w = OracleWorld(sc)
    nd = sc['nd']
    rng = np.random.default_rng(1)
    q = _valid_q(model, rng, nd)
    w.s('S_Q', nd)[:] = q
    free_dynamics(w, sc)
    M, G, _ = mass_matrix_gravity_energy(sc, w, q)
    assert np.allclose(M, M.T) and np.all(np.linalg.eigvalsh(M) > 0)
    tau = M @ rng.uniform(-1, 1, nd)
    w.s('S_JTORQUE', nd)[:] = tau
    w.refresh()
    w.step_physics()
    qdd = w.s('S_QD', nd) / dt
    ref = np.linalg.solve(M, tau - G)
    assert np.abs(qdd - ref).max() <= 1e-6 * np.abs(ref).max()


def test_energy_error_of_undamped_arm_vanishes_with_step_size():
    """Velocity-product (Coriolis/centrifugal) terms: an undamped, unactuated UR5 released from rest conserves
    energy up to the first-order integrator error, which must shrink ~linearly with the step."""
    drifts, kins = [], []
    for hz in (960, 3840):
        sc = build_single('ur5/ur5_robot.urdf', world=dict(timestep=1. / hz, substeps=1))
        w = OracleWorld(sc)
        nd = sc['nd']
        free_dynamics(w, sc)
        q0 = np.array([0.3, -0.4, 0.8, -0.3, 0.5, 0.1])
        w.s('S_Q', nd)[:] = q0
        w.refresh()
        e0 = mass_matrix_gravity_energy(sc, w, q0, np.zeros(nd))[2]
        for _ in range(hz // 4):
            w.step_physics()
        M, _, e1 = mass_matrix_gravity_energy(sc, w, w.s('S_Q', nd).copy(), w.s('S_QD', nd).copy())
        drifts.append(abs(e1 - e0))
        kins.append(0.5 * w.s('S_QD', nd) @ M @ w.s('S_QD', nd))
    assert kins[1] > 5.0  # it really moved
    assert drifts[1] < 0.35 * drifts[0]
    assert drifts[1] < 0.01 * kins[1]


def test_free_fall_closed_form_with_velocity_damping():
    """Sphere in free fall: v_{n+1} = v_n + h (g - k v (1 + |v|)) with k = 0.04 (App. A.2), z_{n+1} = z_n + h v_{n+1}."""
    sc = build_single('sphere2.urdf', xyz=(0, 0, 50.0))
    w = OracleWorld(sc)
    h, v, z = 1 / 480., 0.0, 50.0
    for _ in range(100):
        w.step_physics()
        for _ in range(2):
            v = v + h * (-9.81 - 0.04 * v * (1 + abs(v)))
            z = z + h * v
    assert np.isclose(w.s('S_BVEL', 3)[2], v, rtol=1e-12)
    assert np.isclose(w.s('S_BPOS', 3)[2], z, rtol=1e-12)


def test_torque_free_rotation_conserves_angular_momentum():
    sb = SceneBuilder(gravity=(0, 0, 0))
    sb.add_body('drone', resolve_model('hector_quadrotor/quadrotor.urdf'), xyz=(0, 0, 5), mass=4.0)
    sc = sb.finalize()
    w = OracleWorld(sc)
    free_dynamics(w, sc)
    w.s('S_BOMEGA', 3)[:] = [1.0, 2.0, 3.0]
    w.refresh()

    def ang_mom():
        # composite: base + 4 point masses (motor links have zero inertia), about the world origin
        mass = w.p('P_MASS', sc['nframes'])
        inert = w.p('P_INERTIA', 3 * sc['nframes']).reshape(-1, 3)
        Ltot, ptot = np.zeros(3), np.zeros(3)
        for f in range(sc['nframes']):
            fs = w.frame_state(f)
            R = quat_to_mat(fs['com_quat'])
            Ltot += R @ np.diag(inert[f]) @ R.T @ fs['omega'] + mass[f] * np.cross(fs['com_pos'], fs['vel'])
            ptot += mass[f] * fs['vel']
        return Ltot, ptot

    # the base velocity must be consistent with a pure spin of the composite; let it run and compare momenta
    L0, p0 = ang_mom()
    for _ in range(240):
        w.step_physics()
    L1, p1 = ang_mom()
    assert np.allclose(p1, p0, rtol=2e-3, atol=1e-4)
    assert np.allclose(L1, L0, rtol=2e-3, atol=1e-6)  # first-order integrator: small drift allowed


def test_quadrotor_is_a_five_link_tree_of_mass_8():
    """mass override 4.0 on the base + four inertial-less motor links of mass 1 each (App. A.1)."""
    sb = SceneBuilder()
    sb.add_body('drone', resolve_model('hector_quadrotor/quadrotor.urdf'), xyz=(0, 0, 5), mass=4.0)
    sc = sb.finalize()
    w = OracleWorld(sc)
    assert np.isclose(w.p('P_MASS', 5).sum(), 8.0)
    free_dynamics(w, sc)
    # a world-frame push F on the base for one step accelerates the composite by F / 8
    w.s('S_EXTF', 3)[:] = [0, 0, 8.0 * 9.81 + 8.0]
    w.refresh()
    w.step_physics()
    assert np.isclose(w.s('S_BVEL', 3)[2], 1.0 / 240., rtol=1e-9)


def test_one_dof_motor_impulse_is_clamped():
    """PGS on a single velocity motor: applied impulse = clamp(m_eff (v* - v), +-F dt) (App. A.3)."""
    sc = build_single('ur5/ur5_robot.urdf', world=dict(gravity=(0, 0, 0)))
    w = OracleWorld(sc)
    nd = sc['nd']
    free_dynamics(w, sc)
    w.s('S_MMAXF', nd)[0] = 150.0   # shoulder pan only
    w.s('S_MKD', nd)[0] = 1.0
    w.s('S_MTVEL', nd)[0] = 100.0  # unreachable in one step => saturates
    w.refresh()
    w.step_physics()
    assert np.isclose(w.s('S_MAPPLIED', nd)[0], 150.0 / 240.)
    M = mass_matrix_gravity_energy(sc, w, np.zeros(nd))[0]
    # two substeps, each applying the clamped impulse through M^-1
    qd_expected = 2 * np.linalg.solve(M, np.eye(nd)[0] * 150.0 / 240.)
    assert np.allclose(w.s('S_QD', nd), qd_expected, rtol=1e-2, atol=2e-5)
    # reachable target: the motor hits it exactly and the impulse is m_eff * v*
    w2 = OracleWorld(sc)
    free_dynamics(w2, sc)
    w2.s('S_MMAXF', nd)[5] = 28.0
    w2.s('S_MTVEL', nd)[5] = 0.01
    w2.refresh()
    w2.step_physics()
    assert np.isclose(w2.s('S_QD', nd)[5], 0.01, rtol=1e-9)


def test_sphere_rests_on_plane_with_contact_force_mg():
    sb = SceneBuilder()
    sb.add_body('plane', resolve_model('plane.urdf'))
    sb.add_body('ball', resolve_model('sphere2.urdf'), xyz=(0, 0, 0.5))
    sc = sb.finalize()
    w = OracleWorld(sc)
    w.refresh()
    for _ in range(240):
        w.step_physics()
    assert abs(w.s('S_BPOS', 6)[5] - 0.5) < 1e-3
    assert abs(w.s('S_BVEL', 6)[5]) < 1e-3
    cs = w.contacts()
    assert len(cs) == 1 and cs[0]['fa'] == 0 and np.allclose(cs[0]['n'], [0, 0, -1])  # normal points from B (ball) to A (plane)


def test_joint_limit_pushes_back():
    sc = build_single('jaco/j2s7s300_standalone.urdf', world=dict(gravity=(0, 0, 0)))
    w = OracleWorld(sc)
    nd = sc['nd']
    free_dynamics(w, sc)
    q = np.array([0, 3.0, 0, 3.0, 0, 3.0, 0, 1.0, 1.0, 1.0])
    q[7] = -0.05  # finger 1 below its lower limit 0
    w.s('S_Q', nd)[:] = q
    w.refresh()
    w.step_physics()
    assert w.s('S_QD', nd)[7] > 0
    assert w.s('S_Q', nd)[7] > -0.05


def test_ik_matches_numeric_jacobian_damped_least_squares():
    """The oracle's IK equals an independent numpy DLS (numeric Jacobian from FK, lambda = 0.5 on the diagonal of
    J^T J, 20 iterations).  With that default damping one call is deliberately sluggish near singular postures - the
    well-known reason pybullet users iterate calculateInverseKinematics."""
    sc = build_single('ur5/ur5_robot.urdf', xyz=(0.3, 0.1, 0.0), quat=quat_from_euler([0, 0, 0.7]))
    w = OracleWorld(sc)
    nd = sc['nd']
    q0 = np.array([-0.17, -0.73, -1.93, -0.36, -0.03, -0.06])
    ee = sc.bodies[0].frame(7)

    def fk(q):
        w.s('S_Q', nd)[:] = q
        w.refresh()
        return w.frame_state(ee)['link_pos'].copy()

    target = fk(q0) + np.array([0.03, -0.02, 0.04])
    q = q0.copy()
    res = []
    for _ in range(20):
        e = target - fk(q)
        res.append(np.linalg.norm(e))
        J = np.zeros((3, nd))
        for i in range(nd):
            dq = np.zeros(nd)
            dq[i] = 1e-6
            J[:, i] = (fk(q + dq) - fk(q - dq)) / 2e-6
        q = q + np.linalg.solve(J.T @ J + 0.5 * np.eye(nd), J.T @ e)
    w.s('S_Q', nd)[:] = q0
    w.refresh()
    sol = w.ik(0, 7, target)
    assert np.allclose(sol, q, atol=1e-8)
    assert all(b < a for a, b in zip(res, res[1:]))


def test_euler_quaternion_helpers_round_trip():
    import ctypes
    from oracle.oracle import lib
    rng = np.random.default_rng(5)
    for _ in range(20):
        rpy = rng.uniform([-3, -1.5, -3], [3, 1.5, 3])
        q = quat_from_euler(rpy)
        # Rz Ry Rx convention
        cr, sr, cp, sp, cy, sy = np.cos(rpy[0]), np.sin(rpy[0]), np.cos(rpy[1]), np.sin(rpy[1]), np.cos(rpy[2]), np.sin(rpy[2])
        Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
        Ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
        Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
        assert np.allclose(quat_to_mat(q), Rz @ Ry @ Rx, atol=1e-12)


def test_admittance_gravity_compensation_matches_independent_gravity_vector():
    """admittance_controller.py:47 (p.calculateInverseDynamics(q, 0, 0)): with no commanded wrench and zero gains the op
    writes exactly G(q) into the applied joint torques - checked against the numpy potential-energy gradient of
    tests/helpers.py - and a motor-less, damping-less arm then stays where it is."""
    from diy_gym_b200.assets import resolve_model
    from diy_gym_b200.compiler.scene import SceneBuilder
    sb = SceneBuilder()
    b = sb.add_body('m', resolve_model('ur5/ur5_robot.urdf'))
    nd = b.n_dofs
    ee = b.joint_names().index('ee_fixed_joint')
    dofs = [b.global_dof(i) for i in b.movable_joints()]
    sb.add_op('ADMITTANCE', [b.index, b.link_start + ee, nd] + dofs, [0.0, 0.0, 0.0, 0.0, 0.05] + [0.0] * nd, n_act=6)
    sb.motors_off += dofs
    sc = sb.finalize()
    w = OracleWorld(sc)
    free_dynamics(w, sc)
    q = np.array([0.3, -1.1, 1.4, -0.4, 0.7, 0.2])
    w.s('S_Q', nd)[:] = q
    w.refresh()
    _, G, _ = mass_matrix_gravity_energy(sc, w, q)
    w.apply_actions(np.zeros(6))
    assert np.allclose(w.s('S_JTORQUE', nd), G, rtol=1e-10, atol=1e-10)
    # a pure force along +x at the admittance point adds J_lin^T F: compare with a finite difference of the point position
    w.s('S_JTORQUE', nd)[:] = 0
    w.apply_actions(np.array([1.0, 0, 0, 0, 0, 0]))
    tau_f = w.s('S_JTORQUE', nd) - G
    lpos, lquat = w.s('S_LPOS', 3 * sc['nl']).copy().reshape(-1, 3), w.s('S_LQUAT', 4 * sc['nl']).copy().reshape(-1, 4)
    from diy_gym_b200.compiler.mathutil import quat_rotate
    p0 = lpos[ee] + quat_rotate(lquat[ee], np.array([0, 0, 0.05]))
    for j in range(nd):
        dq = q.copy(); dq[j] += 1e-6
        w.s('S_Q', nd)[:] = dq
        w.refresh()
        lp, lq = w.s('S_LPOS', 3 * sc['nl']).reshape(-1, 3), w.s('S_LQUAT', 4 * sc['nl']).reshape(-1, 4)
        p1 = lp[ee] + quat_rotate(lq[ee], np.array([0, 0, 0.05]))
        assert abs((p1 - p0)[0] / 1e-6 - tau_f[j]) < 1e-5
    # hold still under gravity compensation
    w.s('S_Q', nd)[:] = q
    w.s('S_QD', nd)[:] = 0
    w.refresh()
    for _ in range(50):
        w.s('S_JTORQUE', nd)[:] = 0
        w.apply_actions(np.zeros(6))
        w.step_physics()
    assert np.abs(w.s('S_Q', nd) - q).max() < 1e-6


def test_joint_reaction_wrench_of_a_held_arm_carries_the_weight_above_it():
    """force_torque_sensor.py:21-23: with gravity compensated (zero acceleration) the reaction force of joint j, rotated to the
    world, is the weight of every link from j outwards, and its torque balances their moment about the link's COM."""
    sb = SceneBuilder()
    b = sb.add_body('m', resolve_model('ur5/ur5_robot.urdf'))
    nd = b.n_dofs
    names = b.joint_names()
    ee = names.index('ee_fixed_joint')
    dofs = [b.global_dof(i) for i in b.movable_joints()]
    sb.add_op('ADMITTANCE', [b.index, b.link_start + ee, nd] + dofs, [0.0, 0.0, 0.0, 0.0, 0.0] + [0.0] * nd, n_act=6)
    sb.motors_off += dofs
    sb.need_jreact = True
    j = names.index('wrist_1_joint')
    sb.add_op('FT_SENSOR', [b.link_start + j], n_obs=6)
    sc = sb.finalize()
    w = OracleWorld(sc)
    free_dynamics(w, sc)
    q = np.array([0.3, -1.1, 1.4, -0.4, 0.7, 0.2])
    w.s('S_Q', nd)[:] = q
    w.refresh()
    w.apply_actions(np.zeros(6))
    w.step_physics()
    obs = w.observe()[0]
    nl = sc['nl']
    mass = w.p('P_MASS', sc['nframes'])[sc['nb']:]
    lpos, lquat = w.s('S_LPOS', 3 * nl).reshape(-1, 3), w.s('S_LQUAT', 4 * nl).reshape(-1, 4)
    parent = [b.links[k + 1]['parent'] - 1 for k in range(nl)]
    below = [k for k in range(nl) if any(a == j for a in _ancestors(k, parent))]
    weight = sum(mass[k] for k in below) * 9.81
    R = quat_to_mat(lquat[j])
    f_world = R @ obs[0:3]
    assert np.allclose(f_world, [0, 0, weight], atol=1e-6 * max(weight, 1))
    moment = sum(np.cross(lpos[k] - lpos[j], [0, 0, -mass[k] * 9.81]) for k in below)
    assert np.allclose(R @ obs[3:6], -moment, atol=1e-5)


def _ancestors(k, parent):
    while k >= 0:
        yield k
        k = parent[k]

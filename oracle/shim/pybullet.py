"""TEST INFRASTRUCTURE - a module named `pybullet` that implements exactly the API surface the reference uses
 on top of this repo's fp64 CPU oracle (oracle/bullet_restatement.c), one environment.

Purpose: run the reference's OWN Python layer - /root/reference/diy_gym/{diy_gym,model,config,utils}.py and every
add-on under diy_gym/addons/ - unmodified, and record what it returns (tools/make_reference_layer_golden.py).
Those records pin this repo's host layer and add-on ops against the reference's add-on arithmetic, bookkeeping
(joint selection, slicing, dict structure, ordering) and API usage.  They do NOT pin the physics against the real
pybullet: below this module the numbers come from the oracle ("parity unpinned").

Bodies are collected while the reference constructs its models; the oracle world is built at the first call that
needs simulation state (it is rebuilt, carrying the state over, if a body is loaded later).
"""
import math
import os

import numpy as np

from diy_gym_b200.compiler import urdf as _urdf
from diy_gym_b200.compiler.mathutil import quat_from_euler as _qfe
from diy_gym_b200.compiler.scene import SceneBuilder
from oracle.oracle import OracleWorld

GUI, DIRECT, SHARED_MEMORY = 1, 2, 3
JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_FIXED = 0, 1, 4
POSITION_CONTROL, VELOCITY_CONTROL, TORQUE_CONTROL = 2, 0, 1
WORLD_FRAME, LINK_FRAME = 2, 1
ER_NO_SEGMENTATION_MASK = 4


class error(Exception):
    pass


def _vec3(v):
    """pybullet converts each element with PyFloat_AsDouble, so [0, 0, array([x])] is a valid vector."""
    return np.array([float(np.asarray(e).reshape(-1)[0]) for e in v], float)


class _Sim:
    def __init__(self):
        self.params = dict(timestep=1 / 240., substeps=1, iterations=50, gravity=(0, 0, 0))
        self.specs = []        # per body: dict(desc, scale, fixed, pos, quat, mass)
        self.pending_q = {}    # (uid, joint) -> (q, qd) set before the world exists
        self.constraints = []  # createConstraint(JOINT_FIXED) records
        self.ft_sensors = False
        self.dirty = False     # a spec changed after the world was built: rebuild (state carried over) at the next query
        self.world = None
        self.scene = None

    # ---- building ----
    def builder(self):
        p = self.params
        sb = SceneBuilder(timestep=p['timestep'], substeps=max(int(p['substeps']), 1), iterations=int(p['iterations']), gravity=p['gravity'], hot_start=0)
        for i, s in enumerate(self.specs):
            sb.add_body('body%d' % i, s['desc'], xyz=s['pos'], quat=s['quat'], scale=s['scale'], fixed_base=s['fixed'], mass=s['mass'], color=s.get('color'))
        for c in self.constraints:
            sb.add_fixed_constraint(sb.bodies[c[0]], c[1], sb.bodies[c[2]], c[3], c[4], c[5], c[6], c[7])
        sb.need_jreact = self.ft_sensors
        return sb

    def ensure(self):
        if self.world is not None and self.scene['nb'] == len(self.specs) and self.scene['ncons'] == len(self.constraints) and \
                bool(self.scene.builder.need_jreact) == self.ft_sensors and not self.dirty:
            return self.world
        self.dirty = False
        old, old_scene = self.world, self.scene
        sb = self.builder()
        self.scene = sb.finalize()
        self.world = OracleWorld(self.scene)
        w, h = self.world, self.scene.hdr
        if old is not None:   # carry the state of the bodies that already existed
            nbo, ndo = old_scene['nb'], old_scene['nd']
            for key, width, n in (('S_BPOS', 3, nbo), ('S_BQUAT', 4, nbo), ('S_BVEL', 3, nbo), ('S_BOMEGA', 3, nbo)):
                w.state[h[key]:h[key] + width * n] = old.state[old_scene.hdr[key]:old_scene.hdr[key] + width * n]
            for key in ('S_Q', 'S_QD', 'S_MKP', 'S_MKD', 'S_MTPOS', 'S_MTVEL', 'S_MMAXF'):
                w.state[h[key]:h[key] + ndo] = old.state[old_scene.hdr[key]:old_scene.hdr[key] + ndo]
        for (uid, j), (q, qd) in self.pending_q.items():
            d = self.scene.bodies[uid].global_dof(j)
            w.state[h['S_Q'] + d], w.state[h['S_QD'] + d] = q, qd
        self.pending_q = {}
        for i, s in enumerate(self.specs):
            w.state[h['S_BPOS'] + 3 * i:h['S_BPOS'] + 3 * i + 3] = s['pos']
            w.state[h['S_BQUAT'] + 4 * i:h['S_BQUAT'] + 4 * i + 4] = s['quat']
        w.refresh()
        return w

    def body(self, uid):
        self.ensure()
        return self.scene.bodies[uid]


_sim = _Sim()


# ---------------------------------------------------------------- world ------------------------------------------
def connect(mode, *a, **k):
    return -1 if mode == SHARED_MEMORY else 0


def disconnect(*a, **k):
    pass


def resetDebugVisualizerCamera(*a, **k):
    pass


def resetSimulation(*a, **k):
    global _sim
    _sim = _Sim()


def setPhysicsEngineParameter(numSolverIterations=None, numSubSteps=None, fixedTimeStep=None, **k):
    if numSolverIterations is not None:
        _sim.params['iterations'] = numSolverIterations
    if numSubSteps is not None:
        _sim.params['substeps'] = numSubSteps
    if fixedTimeStep is not None:
        _sim.params['timestep'] = fixedTimeStep


def setGravity(x, y, z, **k):
    _sim.params['gravity'] = (x, y, z)


def stepSimulation(*a, **k):
    _sim.ensure().step_physics()


# ---------------------------------------------------------------- load / edit ------------------------------------
def loadURDF(path, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), useFixedBase=False, globalScaling=1.0, **k):
    desc = _urdf.compile_urdf(path, rel_name=os.path.basename(path))
    _sim.specs.append(dict(desc=desc, scale=float(globalScaling), fixed=bool(useFixedBase), pos=np.array(basePosition, float),
                           quat=np.array(baseOrientation, float), mass=None))
    return len(_sim.specs) - 1


def resetBasePositionAndOrientation(uid, pos, quat, **k):
    s = _sim.specs[uid]
    s['pos'], s['quat'] = np.array(pos, float), np.array(quat, float)
    if _sim.world is not None and _sim.scene['nb'] == len(_sim.specs):
        w, h = _sim.world, _sim.scene.hdr
        w.state[h['S_BPOS'] + 3 * uid:h['S_BPOS'] + 3 * uid + 3] = s['pos']
        w.state[h['S_BQUAT'] + 4 * uid:h['S_BQUAT'] + 4 * uid + 4] = s['quat']
        w.state[h['S_BVEL'] + 3 * uid:h['S_BVEL'] + 3 * uid + 3] = 0      # resetBase... also zeroes the base velocity
        w.state[h['S_BOMEGA'] + 3 * uid:h['S_BOMEGA'] + 3 * uid + 3] = 0
        w.refresh()


def createConstraint(parentBodyUniqueId, parentLinkIndex, childBodyUniqueId, childLinkIndex, jointType, jointAxis, parentFramePosition,
                     childFramePosition, parentFrameOrientation=(0, 0, 0, 1), childFrameOrientation=(0, 0, 0, 1), **k):
    """JOINT_FIXED only (diy_gym/model.py:74-75): the two joint frames, given in the COM frames of the two links."""
    if jointType != JOINT_FIXED:
        raise error('the oracle shim implements JOINT_FIXED constraints only')
    _sim.constraints.append((parentBodyUniqueId, parentLinkIndex, childBodyUniqueId, childLinkIndex, _vec3(parentFramePosition),
                             np.array(parentFrameOrientation, float), _vec3(childFramePosition), np.array(childFrameOrientation, float)))
    return len(_sim.constraints) - 1


def enableJointForceTorqueSensor(uid, joint, enableSensor=True, **k):
    _sim.ft_sensors = _sim.ft_sensors or bool(enableSensor)


def changeDynamics(uid, link, mass=None, angularDamping=None, **k):
    if mass is not None and link == -1:
        _sim.specs[uid]['mass'] = float(mass)
        _sim.dirty = _sim.world is not None   # (the mass is a scene constant of the oracle world: rebuild, state carried over)
        return
    raise error('changeDynamics: only the base mass is supported by the oracle shim')


def changeVisualShape(uid, link=-1, rgbaColor=None, **k):
    """model.py:82-83: the colour of the BASE link's visual shapes (link -1); other arguments (textures) are ignored."""
    if rgbaColor is not None and link == -1:
        _sim.specs[uid]['color'] = [float(c) for c in rgbaColor]
        _sim.dirty = _sim.world is not None


def loadTexture(*a, **k):
    return 0


def resetJointState(uid, joint, targetValue, targetVelocity=0.0, **k):
    if _sim.world is None or _sim.scene['nb'] != len(_sim.specs):
        _sim.pending_q[(uid, joint)] = (float(targetValue), float(targetVelocity))
        return
    w, h = _sim.world, _sim.scene.hdr
    d = _sim.scene.bodies[uid].global_dof(joint)
    w.state[h['S_Q'] + d], w.state[h['S_QD'] + d] = targetValue, targetVelocity
    w.refresh()


# ---------------------------------------------------------------- introspection ----------------------------------
def _desc_body(uid):
    """BodyInfo-like view that works before the world exists."""
    from diy_gym_b200.compiler.scene import BodyInfo
    s = _sim.specs[uid]
    return BodyInfo(uid, 'body%d' % uid, s['desc'], s['scale'], s['fixed'], s['pos'], s['quat'], s['mass'], None)


def getNumJoints(uid, **k):
    return _desc_body(uid).num_joints()


def getJointInfo(uid, i, **k):
    b = _desc_body(uid)
    if i < 0 or i >= b.num_joints():
        raise error('getJointInfo failed.')
    ji = b.joint_info(i)
    jtype = {'revolute': JOINT_REVOLUTE, 'continuous': JOINT_REVOLUTE, 'prismatic': JOINT_PRISMATIC, 'fixed': JOINT_FIXED}[ji['type']]
    q_index = ji['q_index']
    return (i, ji['name'].encode(), jtype, q_index, q_index - 1 if q_index >= 0 else -1, 1, ji['damping'], ji['friction'], ji['lower'], ji['upper'],
            ji['max_force'], ji['max_velocity'], ji['link_name'].encode(), (0.0, 0.0, 1.0), (0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0), ji['parent_index'])


def getDynamicsInfo(uid, link, **k):
    w = _sim.ensure()
    f = _sim.scene.bodies[uid].frame(link)
    m = w.param[_sim.scene.hdr['P_MASS'] + f]
    return (m, 0.5)


# ---------------------------------------------------------------- actuation --------------------------------------
def setJointMotorControlArray(uid, jointIndices, controlMode, targetPositions=None, targetVelocities=None, forces=None,
                              positionGains=None, velocityGains=None, **k):
    w, h, b = _sim.ensure(), _sim.scene.hdr, _sim.scene.bodies[uid]
    n = len(jointIndices)
    for i, j in enumerate(jointIndices):
        d = b.global_dof(j)
        if controlMode == TORQUE_CONTROL:
            w.state[h['S_JTORQUE'] + d] += forces[i]
            continue
        kp = positionGains[i] if positionGains is not None else 0.1
        kd = velocityGains[i] if velocityGains is not None else 1.0
        w.state[h['S_MKD'] + d] = kd
        w.state[h['S_MMAXF'] + d] = forces[i] if forces is not None else 0.0
        if controlMode == POSITION_CONTROL:
            w.state[h['S_MKP'] + d] = kp
            w.state[h['S_MTPOS'] + d] = targetPositions[i]
            w.state[h['S_MTVEL'] + d] = targetVelocities[i] if targetVelocities is not None else 0.0
        else:
            w.state[h['S_MKP'] + d] = 0.0
            w.state[h['S_MTPOS'] + d] = 0.0
            w.state[h['S_MTVEL'] + d] = targetVelocities[i] if targetVelocities is not None else 0.0   # pybullet's default target
    assert n >= 0


def _q_to_mat(q):
    from diy_gym_b200.compiler.mathutil import quat_to_mat
    return quat_to_mat(np.asarray(q, float))


def applyExternalForce(uid, link, forceObj, posObj, flags, **k):
    w, h = _sim.ensure(), _sim.scene.hdr
    f = _sim.scene.bodies[uid].frame(link)
    fs = w.frame_state(f)
    F, P = _vec3(forceObj), _vec3(posObj)
    if flags == LINK_FRAME:
        R = _q_to_mat(fs['com_quat'])
        F, rel = R @ F, R @ P
    else:
        rel = P - fs['com_pos']
    w.state[h['S_EXTF'] + 3 * f:h['S_EXTF'] + 3 * f + 3] += F
    w.state[h['S_EXTT'] + 3 * f:h['S_EXTT'] + 3 * f + 3] += np.cross(rel, F)


def applyExternalTorque(uid, link, torqueObj, flags, **k):
    w, h = _sim.ensure(), _sim.scene.hdr
    f = _sim.scene.bodies[uid].frame(link)
    T = _vec3(torqueObj)
    if flags == LINK_FRAME:
        T = _q_to_mat(w.frame_state(f)['com_quat']) @ T
    w.state[h['S_EXTT'] + 3 * f:h['S_EXTT'] + 3 * f + 3] += T


# ---------------------------------------------------------------- state read -------------------------------------
def getBasePositionAndOrientation(uid, **k):
    if _sim.world is None or _sim.scene['nb'] != len(_sim.specs):
        s = _sim.specs[uid]
        return tuple(s['pos']), tuple(s['quat'])
    fs = _sim.world.frame_state(uid)
    return tuple(fs['com_pos']), tuple(fs['com_quat'])


def getBaseVelocity(uid, **k):
    fs = _sim.ensure().frame_state(uid)
    return tuple(fs['vel']), tuple(fs['omega'])


def getLinkState(uid, link, computeLinkVelocity=0, **k):
    w = _sim.ensure()
    b = _sim.scene.bodies[uid]
    fs = w.frame_state(b.frame(link))
    lf = _sim.scene.sec['LINK_F'][b.link_start + link]
    return (tuple(fs['com_pos']), tuple(fs['com_quat']), tuple(lf[13:16]), tuple(lf[16:20]), tuple(fs['link_pos']), tuple(fs['link_quat']),
            tuple(fs['vel']), tuple(fs['omega']))


def getJointState(uid, joint, **k):
    w, h = _sim.ensure(), _sim.scene.hdr
    d = _sim.scene.bodies[uid].global_dof(joint)
    dt = _sim.scene.hdr_f['dt']
    react = (0.0, ) * 6
    if _sim.ft_sensors:   # [Fx Fy Fz Mx My Mz] of the joint's child link, from the last forward-dynamics pass
        gl = _sim.scene.bodies[uid].link_start + joint
        react = tuple(w.state[h['S_JREACT'] + 6 * gl:h['S_JREACT'] + 6 * gl + 6])
    return (w.state[h['S_Q'] + d], w.state[h['S_QD'] + d], react, w.state[h['S_MAPPLIED'] + d] / dt)


def getJointStates(uid, joints, **k):
    return [getJointState(uid, j) for j in joints]


# ---------------------------------------------------------------- solvers ----------------------------------------
def calculateInverseKinematics(bodyUniqueId, endEffectorLinkIndex, targetPosition, targetOrientation=None, lowerLimits=None, upperLimits=None,
                               jointRanges=None, restPoses=None, **k):
    uid = bodyUniqueId
    w = _sim.ensure()
    b = _sim.scene.bodies[uid]
    nd = b.n_dofs
    null = None
    lists = (lowerLimits, upperLimits, jointRanges, restPoses)
    if all(v is not None and len(v) == nd for v in lists):   # null-space terms only when every list spans all DoF
        null = tuple(np.array(v, float) for v in lists)
    sol = w.ik(uid, b.link_start + endEffectorLinkIndex, np.array(targetPosition, float),
               None if targetOrientation is None else np.array(targetOrientation, float), null)
    return tuple(sol)


def _joint_axes_at(uid, q):
    """World axis / origin of every joint and world COM of every link of body uid at joint coordinates q (plain product of
    the scene's link records; numpy, independent of the C code paths it is used to check)."""
    b, sc = _sim.scene.bodies[uid], _sim.scene
    w, h = _sim.ensure(), sc.hdr
    LI, LF = sc.sec['LINK_I'], sc.sec['LINK_F']
    R = {-1: _q_to_mat(w.state[h['S_BQUAT'] + 4 * uid:h['S_BQUAT'] + 4 * uid + 4])}
    P = {-1: np.array(w.state[h['S_BPOS'] + 3 * uid:h['S_BPOS'] + 3 * uid + 3])}
    axes, orgs = {}, {}
    for k in range(b.n_links):
        gl = b.link_start + k
        li, lf = LI[gl], LF[gl]
        par = -1 if li[1] < 0 else li[1] - b.link_start
        R0, a, d = _q_to_mat(lf[0:4]), np.array(lf[10:13]), np.array(lf[7:10])
        qk = q[b.joint_dof[k]] if b.joint_dof[k] >= 0 else 0.0
        if li[2] == 1:
            K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
            Rrel = R0 @ (np.eye(3) + np.sin(qk) * K + (1 - np.cos(qk)) * K @ K)
            r = np.array(lf[4:7]) + Rrel @ d
        elif li[2] == 2:
            Rrel = R0
            r = np.array(lf[4:7]) + R0 @ (d + a * qk)
        else:
            Rrel = R0
            r = np.array(lf[4:7]) + R0 @ d
        R[k] = R[par] @ Rrel
        P[k] = P[par] + R[par] @ r
        axes[k] = R[k] @ a
        orgs[k] = P[k] - R[k] @ d
    return b, LI, R, P, axes, orgs


def calculateJacobian(bodyUniqueId, linkIndex, localPosition, objPositions, objVelocities, objAccelerations, **k):
    """Translational / rotational Jacobian (3 x nDoF each, world axes) of the point `localPosition` of link `linkIndex`, given
    in that link's COM frame (admittance_controller.py:39-45).  Fixed-base bodies."""
    b, LI, R, P, axes, orgs = _joint_axes_at(bodyUniqueId, list(objPositions))
    if len(objPositions) != b.n_dofs:
        raise error('calculateJacobian: %d joint positions for a body with %d degrees of freedom' % (len(objPositions), b.n_dofs))
    pt = P[linkIndex] + R[linkIndex] @ _vec3(localPosition)
    Jl, Ja = np.zeros((3, b.n_dofs)), np.zeros((3, b.n_dofs))
    kk = linkIndex
    while kk >= 0:
        d = b.joint_dof[kk]
        jt = LI[b.link_start + kk][2]
        if d >= 0:
            if jt == 1:
                Jl[:, d], Ja[:, d] = np.cross(axes[kk], pt - orgs[kk]), axes[kk]
            else:
                Jl[:, d] = axes[kk]
        par = LI[b.link_start + kk][1]
        kk = -1 if par < 0 else par - b.link_start
    return tuple(map(tuple, Jl)), tuple(map(tuple, Ja))


def calculateInverseDynamics(bodyUniqueId, objPositions, objVelocities, objAccelerations, **k):
    """Generalized forces that hold the body at objPositions; the reference only calls it with zero velocities and
    accelerations (admittance_controller.py:47), i.e. asks for the gravity torques."""
    if any(abs(v) > 0 for v in objVelocities) or any(abs(v) > 0 for v in objAccelerations):
        raise error('the oracle shim implements calculateInverseDynamics at zero velocity / acceleration only')
    b, LI, R, P, axes, orgs = _joint_axes_at(bodyUniqueId, list(objPositions))
    if len(objPositions) != b.n_dofs:
        raise error('calculateInverseDynamics: %d joint positions for a body with %d degrees of freedom' % (len(objPositions), b.n_dofs))
    w, h = _sim.ensure(), _sim.scene.hdr
    g = np.array(_sim.params['gravity'], float)
    tau = np.zeros(b.n_dofs)
    for k2 in range(b.n_links):
        m = w.param[h['P_MASS'] + b.frame(k2)]
        kk = k2
        while kk >= 0:
            d = b.joint_dof[kk]
            if d >= 0:
                dp = np.cross(axes[kk], P[k2] - orgs[kk]) if LI[b.link_start + kk][2] == 1 else axes[kk]
                tau[d] -= m * g.dot(dp)
            par = LI[b.link_start + kk][1]
            kk = -1 if par < 0 else par - b.link_start
    return tuple(tau)


# ---------------------------------------------------------------- math helpers -----------------------------------
def getQuaternionFromEuler(rpy, **k):
    return tuple(_qfe(rpy))


def getEulerFromQuaternion(q, **k):
    x, y, z, w = [float(v) for v in q]
    sarg = -2 * (x * z - w * y)
    if sarg <= -0.99999:
        return (0.0, -0.5 * math.pi, 2 * math.atan2(x, -y))
    if sarg >= 0.99999:
        return (0.0, 0.5 * math.pi, 2 * math.atan2(-x, y))
    return (math.atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z), math.asin(sarg), math.atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z))


def getMatrixFromQuaternion(q, **k):
    return tuple(_q_to_mat(q).reshape(-1))


def multiplyTransforms(pa, qa, pb, qb, **k):
    from diy_gym_b200.compiler.mathutil import quat_mul, quat_rotate
    return tuple(np.asarray(pa, float) + quat_rotate(np.asarray(qa, float), np.asarray(pb, float))), tuple(quat_mul(np.asarray(qa, float), np.asarray(qb, float)))


def invertTransform(p, q, **k):
    from diy_gym_b200.compiler.mathutil import quat_conj, quat_rotate
    qi = quat_conj(np.asarray(q, float))
    return tuple(-quat_rotate(qi, np.asarray(p, float))), tuple(qi)


def computeProjectionMatrixFOV(fov, aspect, nearVal, farVal, **k):
    f = 1.0 / math.tan(math.radians(fov) / 2)
    n, fa = nearVal, farVal
    return (f / aspect, 0, 0, 0, 0, f, 0, 0, 0, 0, (n + fa) / (n - fa), -1, 0, 0, 2 * fa * n / (n - fa), 0)


def getCameraImage(width, height, viewMatrix=None, projectionMatrix=None, flags=0, **k):
    """(width, height, rgba u8 [H,W,4], depth buffer [H,W] in [0,1], body ids [H,W]) from the two matrices the caller hands over
    - oracle.dgo_get_camera_image consults nothing else, so the reference's own camera conventions decide the image."""
    w = _sim.ensure()
    rgba, depth, segm = w.get_camera_image(width, height, viewMatrix, projectionMatrix)
    return (width, height, rgba, depth, segm)


def getKeyboardEvents(*a, **k):
    return {}

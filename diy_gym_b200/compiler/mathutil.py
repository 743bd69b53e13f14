"""fp64 numpy rigid-transform helpers (quaternions are xyzw, as in the pybullet API the reference uses)."""
import numpy as np


def quat_from_euler(rpy):
    """R = Rz(yaw) Ry(pitch) Rx(roll)  (pybullet `getQuaternionFromEuler`, used at `diy_gym/model.py:54`)."""
    r, p, y = [float(v) for v in rpy]
    cr, sr = np.cos(r / 2), np.sin(r / 2)
    cp, sp = np.cos(p / 2), np.sin(p / 2)
    cy, sy = np.cos(y / 2), np.sin(y / 2)
    return np.array([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                     cr * cp * cy + sr * sp * sy])


def quat_mul(a, b):
    """Hamilton product a*b (rotation b applied first)."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def quat_conj(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


def quat_to_mat(q):
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def mat_to_quat(m):
    t = np.trace(m)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [(m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s]
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        q = [0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s, (m[2, 1] - m[1, 2]) / s]
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        q = [(m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s, (m[0, 2] - m[2, 0]) / s]
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        q = [(m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s, (m[1, 0] - m[0, 1]) / s]
    q = np.array(q)
    return q / np.linalg.norm(q)


def quat_rotate(q, v):
    return quat_to_mat(q) @ np.asarray(v, dtype=float)


class Transform:
    """Rigid transform  x_parent = R x_child + p  stored as (p, q)."""
    __slots__ = ('p', 'q')

    def __init__(self, p=(0, 0, 0), q=(0, 0, 0, 1)):
        self.p = np.asarray(p, dtype=float).copy()
        self.q = np.asarray(q, dtype=float).copy()

    @classmethod
    def from_xyz_rpy(cls, xyz, rpy):
        return cls(xyz, quat_from_euler(rpy))

    def __mul__(self, o):
        return Transform(self.p + quat_rotate(self.q, o.p), quat_mul(self.q, o.q))

    def inverse(self):
        qi = quat_conj(self.q)
        return Transform(-quat_rotate(qi, self.p), qi)

    def apply(self, v):
        return self.p + quat_rotate(self.q, v)

    def matrix(self):
        T = np.eye(4)
        T[:3, :3] = quat_to_mat(self.q)
        T[:3, 3] = self.p
        return T

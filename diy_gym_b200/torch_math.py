"""Batched quaternion / vector helpers on torch tensors (xyzw quaternions) for user add-ons and model queries."""
import torch


def quat_mul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz], -1)


def quat_conj(q):
    return q * q.new_tensor([-1.0, -1.0, -1.0, 1.0])


def quat_rotate(q, v):
    """Rotate vectors v [...,3] by unit quaternions q [...,4]."""
    qv, qw = q[..., :3], q[..., 3:4]
    t = 2.0 * torch.cross(qv, v.expand_as(qv) if v.dim() < qv.dim() or v.shape != qv.shape else v, dim=-1)
    vv = v.expand_as(qv) if v.shape != qv.shape else v
    return vv + qw * t + torch.cross(qv, t, dim=-1)


def quat_rotate_inv(q, v):
    return quat_rotate(quat_conj(q), v)


def quat_from_euler(rpy):
    """R = Rz(yaw) Ry(pitch) Rx(roll)  (pybullet getQuaternionFromEuler)."""
    r, p, y = (0.5 * rpy).unbind(-1)
    cr, sr, cp, sp, cy, sy = torch.cos(r), torch.sin(r), torch.cos(p), torch.sin(p), torch.cos(y), torch.sin(y)
    return torch.stack([sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
                        cr * cp * cy + sr * sp * sy], -1)


def euler_from_quat(q):
    """pybullet getEulerFromQuaternion (regular branch; gimbal lock handled as pybullet does)."""
    x, y, z, w = q.unbind(-1)
    sarg = -2 * (x * z - w * y)
    roll = torch.atan2(2 * (y * z + w * x), w * w - x * x - y * y + z * z)
    pitch = torch.asin(sarg.clamp(-1, 1))
    yaw = torch.atan2(2 * (x * y + w * z), w * w + x * x - y * y - z * z)
    lo, hi = sarg <= -0.99999, sarg >= 0.99999
    roll = torch.where(lo | hi, torch.zeros_like(roll), roll)
    pitch = torch.where(lo, torch.full_like(pitch, -0.5 * torch.pi), torch.where(hi, torch.full_like(pitch, 0.5 * torch.pi), pitch))
    yaw = torch.where(lo, 2 * torch.atan2(x, -y), torch.where(hi, 2 * torch.atan2(-x, y), yaw))
    return torch.stack([roll, pitch, yaw], -1)

"""Quaternion / Euler leaf math of the TACO hot path, torch-float32 CPU restatement.
TEST INFRASTRUCTURE (see oracle/__init__.py).

Quaternions are xyzw.  The operation ORDER of each formula follows the reference so that
float32 rounding matches it; the golden vectors in tests/golden/leaf_math.npz were
produced by the reference functions themselves (oracle/make_golden.py).

Reference: python/isaacgym/torch_utils.py (TU) and
IsaacGymEnvs/isaacgymenvs/utils/torch_jit_utils.py (JIT).
"""
import math

import torch

HALF_PI = math.pi / 2.0


def qmul(a, b):
    """Hamilton product, 8-multiply factored form.  TU:19-40 (quat_mul)."""
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    t_ww = (az + ax) * (bx + by)
    t_yy = (aw - ay) * (bw + bz)
    t_zz = (aw + ay) * (bw - bz)
    t_xx = t_ww + t_yy + t_zz
    t_qq = 0.5 * (t_xx + (az - ax) * (bx - by))
    ow = t_qq - t_ww + (az - ay) * (by - bz)
    ox = t_qq - t_xx + (ax + aw) * (bx + bw)
    oy = t_qq - t_yy + (aw - ax) * (by + bz)
    oz = t_qq - t_zz + (az + ay) * (bw - bx)
    return torch.stack((ox, oy, oz, ow), dim=-1)


def qconj(q):
    """TU:84-88 (quat_conjugate)."""
    return torch.cat((-q[..., :3], q[..., 3:]), dim=-1)


def cross3(a, b):
    ax, ay, az = a.unbind(-1)
    bx, by, bz = b.unbind(-1)
    return torch.stack((ay * bz - az * by, az * bx - ax * bz, ax * by - ay * bx), dim=-1)


def qrot(q, v):
    """Rotate v by q:  v(2w^2-1) + 2w(q x v) + 2q(q.v).  TU:58-68 (quat_rotate).

    The reference evaluates q.v with torch.bmm (TU:66-67); a plain sequential
    multiply-add is used here, which differs from a BLAS dot by at most 1 ulp of the
    largest product."""
    w = q[..., 3:4]
    u = q[..., :3]
    part_a = v * (2.0 * w * w - 1.0)
    part_b = cross3(u, v) * w * 2.0
    dot = (u[..., 0:1] * v[..., 0:1] + u[..., 1:2] * v[..., 1:2]) + u[..., 2:3] * v[..., 2:3]
    part_c = u * dot * 2.0
    return part_a + part_b + part_c


def euler_xyz(q):
    """roll, pitch, yaw in (-pi, pi].  TU:175-196 (get_euler_xyz_v1)."""
    x, y, z, w = q.unbind(-1)
    sinr_cosp = 2.0 * (w * x + y * z)
    cosr_cosp = w * w - x * x - y * y + z * z
    roll = torch.atan2(sinr_cosp, cosr_cosp)
    sinp = 2.0 * (w * y - z * x)
    pitch = torch.where(sinp.abs() >= 1, torch.copysign(torch.full_like(sinp, HALF_PI), sinp), torch.asin(sinp))
    siny_cosp = 2.0 * (w * z + x * y)
    cosy_cosp = w * w + x * x - y * y - z * z
    yaw = torch.atan2(siny_cosp, cosy_cosp)
    return roll, pitch, yaw


def roll_of(q):
    """Only the roll component of euler_xyz (the only one the hot path consumes)."""
    x, y, z, w = q.unbind(-1)
    return torch.atan2(2.0 * (w * x + y * z), w * w - x * x - y * y + z * z)


_SIN_C = (-1.0 / 6, 1.0 / 120, -1.0 / 5040, 1.0 / 362880, -1.0 / 39916800, 1.0 / 6227020800)
_COS_C = (-1.0 / 2, 1.0 / 24, -1.0 / 720, 1.0 / 40320, -1.0 / 3628800, 1.0 / 479001600, -1.0 / 87178291200)


def sincos_draw(x):
    """sin(x), cos(x) for |x| <= pi/2 by Horner evaluation of the degree-13 / degree-14 Taylor
    polynomials (truncation < 7e-10 relative).  Used ONLY for the random attitude draws of a reset
    (copter attitude, target yaw, observation-noise quaternion), which are our own Philox-driven
    draws: built from + and * alone, the -fmad=false CUDA kernel reproduces it bit for bit, so
    oracle and kernel start every episode from identical float32 quaternions (libm sin/cos differ
    by 1 ulp between vendors, which decorrelates the subsequent float32 rounding)."""
    z = x * x
    s = torch.full_like(x, _SIN_C[-1])
    for c in reversed(_SIN_C[:-1]):
        s = c + z * s
    s = x * (1.0 + z * s)
    c_ = torch.full_like(x, _COS_C[-1])
    for c in reversed(_COS_C[:-1]):
        c_ = c + z * c_
    c_ = 1.0 + z * c_
    return s, c_


def quat_from_euler(roll, pitch, yaw, sincos=None):
    """TU:199-213 (quat_from_euler_xyz).  ``sincos`` replaces torch.sin/torch.cos of the half
    angles (see sincos_draw); default = torch, as in the reference."""
    if sincos is None:
        sincos = lambda a: (torch.sin(a), torch.cos(a))
    sy, cy = sincos(yaw * 0.5)
    sr, cr = sincos(roll * 0.5)
    sp, cp = sincos(pitch * 0.5)
    qw = cy * cr * cp + sy * sr * sp
    qx = cy * sr * cp - sy * cr * sp
    qy = cy * cr * sp + sy * sr * cp
    qz = sy * cr * cp - cy * sr * sp
    return torch.stack((qx, qy, qz, qw), dim=-1)


def rotmat9(q):
    """Row-major 3x3 rotation matrix of an (un-normalised) xyzw quaternion, flattened
    to 9.  JIT:389-416 (quaternion_to_matrix)."""
    i, j, k, r = q.unbind(-1)
    two_s = 2.0 / (q * q).sum(-1)
    return torch.stack((
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), dim=-1)


def quat_angle(a, b):
    """2*asin(min(|vec(a * conj(b))|, 1)).  JIT:146-164 (quat_diff_rad)."""
    m = qmul(a, qconj(b))
    return 2.0 * torch.asin(torch.clamp(torch.norm(m[..., 0:3], p=2, dim=-1), max=1.0))


def rand_range(lo, hi, u):
    """(hi-lo)*u + lo with python-float bounds.  TU:216-219 (torch_rand_float)."""
    return (hi - lo) * u + lo

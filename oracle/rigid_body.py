"""Free rigid-body step that stands in for ``gym.simulate`` (PhysX) on the TACO path.
TEST INFRASTRUCTURE.

The reference advances the MAV with closed-source PhysX (vec_task_asymmetry.py:313);
the binaries are absent, so this integrator is OUR specification, frozen here and in
DESIGN.md, and "parity" for this row means CUDA kernel == this restatement.

Model facts taken from the reference:
  * mass 0.46 kg chassis + 8 x 1e-7 kg arm/rotor bodies, chassis inertia
    diag(5e-4, 7e-4, 8e-4) kg m^2, rotor hubs at (+-0.047, +-0.059, 0.02) m
    (IsaacGymEnvs/assets/xml/fpv_without_duct.xml:4-43)
  * gravity (0,0,-9.81) (fpv_asymmetry.py:215-217); no damping, no velocity limits
    (fpv_asymmetry.py:250-256); gyroscopic forces on (docs release notes: enabled by default)
  * forces/torques are given in the BODY frame and applied with LOCAL_SPACE
    (fpv_asymmetry.py:620-635); the simulate() call runs ``substeps`` sub-steps of
    h = dt/substeps (vec_task_asymmetry.py:432).
  * root state = pos, quat xyzw, linear velocity (world), angular velocity (world)
    (docs/programming/tensors: actor root state layout).

Our specification (semi-implicit Euler, body-frame rotational dynamics; every line is float32, `fma(a,b,c)` = a*b+c with
ONE rounding, every other operation rounds separately):
    u = q.xyz;  t = 2 (u x F_b);  F_w = (fma(q.w, t, F_b)) + u x t      once per simulate call: the FORCE is rotated once and
                                          held constant in the world frame over the sub-steps (what PhysX does with an applied force)
    a = (F_w.x / m, F_w.y / m, fma(F_w.z, 1/m, -9.81))
    per sub-step:
        v   = fma(h, a, v)
        w_b = fma(h, (tau_b - w_b x (I w_b)) * I^-1, w_b)   the TORQUE is held constant in the BODY frame (rotor thrust
                                          differentials and reaction torques are body-fixed; PhysX would hold it in the world
                                          frame -- an O(h^2) difference at h = 0.5 ms)
        p   = fma(h, v, p)
        q   = renorm(q * exp(h/2 * w_b))                    body-frame increment, right-multiplied (Hamilton product, fma chains)
    a x b = (fma(a.y, b.z, -(a.z b.y)), ...)
    exp(h/2 w_b) ~ ( w_b * (h/2) fma(t, fma(t, 1/120, -1/6), 1),  fma(t, fma(t, 1/24, -1/2), 1) ),  t = |w_b|^2 (h/2)^2
        4th-order series: truncation theta^6/720 < 1e-13 for |w_b| <= 80 rad/s, needs only +, * and fma
    renorm(q) = q * fma(-0.5, |q|^2, 1.5) one Newton step of 1/sqrt about 1 (|q|^2 = 1 + O(1e-7)): no sqrt, no division
The exact order of operations is the C code in oracle/rigid_body.c (the executable form of this specification;
`_integrate_py` below is its pure-numpy twin with an exact float32 FMA emulation, used when the C library is not built and
checked against it bit for bit in tests/test_oracle_semantics.py).  The FMAs are part of the specification because the
integrator is the one part of the path that is OURS: they cut the integrator's instruction count on the GPU by 40 % (9 % of
the whole step) while staying bit-reproducible between the CUDA kernel (explicit __fmaf_rn in a -fmad=false build) and the CPU.
The angular velocity is carried in BODY coordinates (w_b) for the whole VecTask.step; the world-frame
root-state value w = R(q) w_b is materialised by the caller at the end of the RL step and converted back
with w_b = R(q)^T w at the start of the next one (fpv_asymmetry.py:350), so the flip-reset quirk that leaves
the world-frame w_y, w_z stale (fpv_asymmetry.py:876) keeps its meaning.
"""
import ctypes
import os

import numpy as np
import torch

from .leaf_math import qmul, qconj, qrot, cross3

MASS = 0.46 + 8 * 1e-7
INERTIA = (5e-4, 7e-4, 8e-4)
GRAVITY_Z = -9.81
ARM_X, ARM_Y = 0.047, 0.059


def body_wrench(rotor_force_sim, rotor_torque_sim, body_force):
    """Net body-frame force and torque from the four sim-rotor thrusts (along +z at the
    hub positions (+x+y, -x+y, -x-y, +x-y), fpv_without_duct.xml:8-40), the rotor reaction
    torques about z and the chassis drag force (fpv_asymmetry.py:620-627)."""
    f0, f1, f2, f3 = rotor_force_sim.unbind(-1)
    force = torch.stack((body_force[:, 0], body_force[:, 1], body_force[:, 2] + (((f0 + f1) + f2) + f3)), dim=1)
    tx = ARM_Y * (((f0 + f1) - f2) - f3)
    ty = ARM_X * (((f1 + f2) - f0) - f3)
    tz = ((rotor_torque_sim[:, 0] + rotor_torque_sim[:, 1]) + rotor_torque_sim[:, 2]) + rotor_torque_sim[:, 3]
    return force, torch.stack((tx, ty, tz), dim=1)


_LIB = None


def _c_lib():
    """oracle/librigid_body.so (oracle/build_c.py); built on demand when gcc is available, else None."""
    global _LIB
    if _LIB is None:
        here = os.path.dirname(os.path.abspath(__file__))
        path, src = os.path.join(here, "librigid_body.so"), os.path.join(here, "rigid_body.c")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
            try:
                from . import build_c
                build_c.build()
            except Exception:
                pass
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            f = ctypes.c_float
            lib.taco_oracle_integrate.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 6 + [f] * 7 + [ctypes.c_int]
            lib.taco_oracle_integrate.restype = None
            _LIB = lib
        else:
            _LIB = False
    return _LIB or None


def fma32(a, b, c):
    """Exact float32 fused multiply-add on numpy arrays: round(a*b + c) with ONE rounding.  The product of two float32
    is exact in float64; the float64 sum is rounded once more when cast to float32, which can only go wrong when the
    float64 sum lands exactly on a float32 rounding boundary -- then the exact error of the float64 addition (TwoSum)
    decides the direction."""
    a64, b64, c64 = np.asarray(a, np.float64), np.asarray(b, np.float64), np.asarray(c, np.float64)
    p = a64 * b64
    s = p + c64
    bb = s - p
    err = (p - (s - bb)) + (c64 - bb)                    # s + err == a*b + c exactly
    r = s.astype(np.float32)
    rd = r.astype(np.float64)
    diff = s - rd
    other = np.nextafter(r, np.where(diff > 0, np.float32(np.inf), np.float32(-np.inf)).astype(np.float32))
    tie = (diff != 0) & ((rd + other.astype(np.float64)) * 0.5 == s) & (err != 0)
    if np.any(tie):
        up = np.maximum(r, other)
        dn = np.minimum(r, other)
        r = np.where(tie, np.where(err > 0, up, dn), r)
    return r.astype(np.float32)


def _cross_f(a, b):
    return (fma32(a[1], b[2], -(a[2] * b[1])), fma32(a[2], b[0], -(a[0] * b[2])), fma32(a[0], b[1], -(a[1] * b[0])))


def _integrate_py(pos, quat, vel, wb, fb, tb, h, half_h, half_h2, c3, c5, c4, inv_mass, substeps):
    """Pure-numpy twin of oracle/rigid_body.c (same operations in the same order, exact FMA emulation)."""
    f32 = np.float32
    qx, qy, qz, qw = (quat[:, k].copy() for k in range(4))
    p = [pos[:, k].copy() for k in range(3)]
    v = [vel[:, k].copy() for k in range(3)]
    w = [wb[:, k].copy() for k in range(3)]
    f = [fb[:, k] for k in range(3)]
    tq = [tb[:, k] for k in range(3)]
    u = (qx, qy, qz)
    t = _cross_f(u, f)
    t = tuple(x + x for x in t)
    ut = _cross_f(u, t)
    fw = [fma32(qw, t[k], f[k]) + ut[k] for k in range(3)]
    acc = [fw[0] * inv_mass, fw[1] * inv_mass, fma32(fw[2], inv_mass, f32(-9.81))]
    inv_i = (f32(2000.0), f32(1.0 / 7e-4), f32(1250.0))
    ine = (f32(5e-4), f32(7e-4), f32(8e-4))
    for _ in range(substeps):
        v = [fma32(h, acc[k], v[k]) for k in range(3)]
        iw = [ine[k] * w[k] for k in range(3)]
        gy = _cross_f(w, iw)
        w = [fma32(h, (tq[k] - gy[k]) * inv_i[k], w[k]) for k in range(3)]
        p = [fma32(h, v[k], p[k]) for k in range(3)]
        w2 = fma32(w[2], w[2], fma32(w[1], w[1], w[0] * w[0]))
        th2 = w2 * half_h2
        kk = half_h * fma32(th2, fma32(th2, c5, c3), f32(1.0))
        cs = fma32(th2, fma32(th2, c4, f32(-0.5)), f32(1.0))
        dx, dy, dz = w[0] * kk, w[1] * kk, w[2] * kk
        rw = fma32(-qz, dz, fma32(-qy, dy, fma32(-qx, dx, qw * cs)))
        rx = fma32(-qz, dy, fma32(qy, dz, fma32(qx, cs, qw * dx)))
        ry = fma32(-qx, dz, fma32(qz, dx, fma32(qy, cs, qw * dy)))
        rz = fma32(-qy, dx, fma32(qx, dy, fma32(qz, cs, qw * dz)))
        n2 = fma32(rw, rw, fma32(rz, rz, fma32(ry, ry, rx * rx)))
        rn = fma32(f32(-0.5), n2, f32(1.5))
        qx, qy, qz, qw = rx * rn, ry * rn, rz * rn, rw * rn
    return (np.stack(p, 1), np.stack((qx, qy, qz, qw), 1), np.stack(v, 1), np.stack(w, 1))


def integrate(pos, quat, linvel, w_b, force_b, torque_b, dt, substeps, use_c=True):
    """One simulate(dt) call.  All tensors (N,3|4) float32; w_b is the body-frame angular velocity.
    Returns the new (pos, quat, linvel, w_b)."""
    f32 = np.float32
    h64 = float(dt) / substeps                                 # python evaluates these in double; the tensor op rounds them
    hh64 = 0.5 * h64
    consts = (f32(h64), f32(hh64), f32(hh64 * hh64), f32(-1.0 / 6.0), f32(1.0 / 120.0), f32(1.0 / 24.0), f32(1.0 / MASS))
    arrs = [np.ascontiguousarray(t.detach().numpy(), dtype=np.float32).copy() for t in (pos, quat, linvel, w_b)]
    fb = np.ascontiguousarray(force_b.detach().numpy(), dtype=np.float32)
    tb = np.ascontiguousarray(torque_b.detach().numpy(), dtype=np.float32)
    lib = _c_lib() if use_c else None
    if lib is not None:
        n = arrs[0].shape[0]
        lib.taco_oracle_integrate(n, *[a.ctypes.data for a in arrs], fb.ctypes.data, tb.ctypes.data, *[float(c) for c in consts], int(substeps))
        out = arrs
    else:
        with np.errstate(all="ignore"):
            out = _integrate_py(arrs[0], arrs[1], arrs[2], arrs[3], fb, tb, *consts, int(substeps))
    return tuple(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)) for a in out)

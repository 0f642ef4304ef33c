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
    (fpv_asymmetry.py:620-635): converted to the world frame with the pose at apply time
    and held constant for the whole simulate() call, over ``substeps`` sub-steps of
    h = dt/substeps (vec_task_asymmetry.py:432).
  * root state = pos, quat xyzw, linear velocity (world), angular velocity (world)
    (docs/programming/tensors: actor root state layout).

Step (semi-implicit Euler; quaternion exponential by its 4th-order series -- with
theta = |w_b| h/2 <= 0.02 rad for |w_b| <= 80 rad/s the truncation error theta^6/720 is < 1e-13,
far below float32 resolution, and the form needs only + and *, so a -fmad=false CUDA build
reproduces the torch float32 result bit for bit; no sqrt / sin / cos / division):
    exp(h/2 w_b) ~ ( w_b * (h/2)(1 - t/6 + t^2/120),  1 - t/2 + t^2/24 ),  t = theta^2
    F_w = R(q) F_b ; tau_w = R(q) tau_b ; w_b = R(q)^T w        (once per simulate call)
    per sub-step s:
        v    += h * (F_w / m + g)
        tau_s = tau_b if s == 0 else R(q)^T tau_w
        w_b  += h * I^-1 (tau_s - w_b x (I w_b))
        p    += h * v
        q     = renorm(q * exp(h/2 * w_b))      (body-frame increment, right-multiplied;
                                                 it rotates about w_b, so w_b is unchanged)
    w = R(q) w_b                                                  (once, at the end)
    renorm(q) = q * (1.5 - 0.5 |q|^2)    (one Newton step of 1/sqrt: no sqrt, no division)
"""
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


def integrate(pos, quat, linvel, angvel, force_b, torque_b, dt, substeps):
    """One simulate(dt) call.  All tensors (N,3|4) float32; returns the new root state."""
    inertia = torch.tensor(INERTIA, dtype=torch.float32)
    inv_inertia = torch.tensor([1.0 / i for i in INERTIA], dtype=torch.float32)
    inv_mass = 1.0 / MASS
    h = dt / substeps
    grav = torch.tensor([0.0, 0.0, GRAVITY_Z], dtype=torch.float32)
    force_w = qrot(quat, force_b)
    torque_w = qrot(quat, torque_b)
    w_b = qrot(qconj(quat), angvel)
    for s in range(substeps):
        linvel = linvel + h * (force_w * inv_mass + grav)
        tau_b = torque_b if s == 0 else qrot(qconj(quat), torque_w)
        w_b = w_b + h * ((tau_b - cross3(w_b, inertia * w_b)) * inv_inertia)
        pos = pos + h * linvel
        hh = 0.5 * h
        w2 = (w_b[:, 0:1] * w_b[:, 0:1] + w_b[:, 1:2] * w_b[:, 1:2]) + w_b[:, 2:3] * w_b[:, 2:3]
        t = w2 * (hh * hh)                                   # theta^2
        k = hh * (1.0 + t * (-1.0 / 6.0 + t * (1.0 / 120.0)))  # sin(theta)/|w|
        c = 1.0 + t * (-0.5 + t * (1.0 / 24.0))                # cos(theta)
        dq = torch.cat((w_b * k, c), dim=1)
        quat = qmul(quat, dq)
        # renormalise with one Newton step of 1/sqrt about 1: |q|^2 = 1 + e with |e| ~ 1e-7 after a
        # float32 product of unit quaternions, so q * (1.5 - 0.5 |q|^2) is unit to O(e^2) ~ 1e-14
        n2 = ((quat[:, 0:1] * quat[:, 0:1] + quat[:, 1:2] * quat[:, 1:2]) + quat[:, 2:3] * quat[:, 2:3]) + quat[:, 3:4] * quat[:, 3:4]
        quat = quat * (1.5 - 0.5 * n2)
    angvel = qrot(quat, w_b)
    return pos, quat, linvel, angvel

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

Our specification (semi-implicit Euler, body-frame rotational dynamics):
    F_w = R(q) F_b                        once per simulate call: the FORCE is held constant in the world
                                          frame over the sub-steps (what PhysX does with an applied force)
    per sub-step:
        v   += h * (F_w / m + g)
        w_b += h * I^-1 (tau_b - w_b x (I w_b))      the TORQUE is held constant in the BODY frame (rotor
                                          thrust differentials and reaction torques are body-fixed; PhysX would
                                          hold it in the world frame -- an O(h^2) difference at h = 0.5 ms)
        p   += h * v
        q    = renorm(q * exp(h/2 * w_b))            body-frame increment, right-multiplied
    exp(h/2 w_b) ~ ( w_b * (h/2)(1 - t/6 + t^2/120),  1 - t/2 + t^2/24 ),  t = |w_b|^2 (h/2)^2
        4th-order series: truncation theta^6/720 < 1e-13 for |w_b| <= 80 rad/s, needs only + and *
    renorm(q) = q * (1.5 - 0.5 |q|^2)     one Newton step of 1/sqrt about 1 (|q|^2 = 1 + O(1e-7)): no sqrt, no division
The angular velocity is carried in BODY coordinates (w_b) for the whole VecTask.step; the world-frame
root-state value w = R(q) w_b is materialised by the caller at the end of the RL step and converted back
with w_b = R(q)^T w at the start of the next one (fpv_asymmetry.py:350), so the flip-reset quirk that leaves
the world-frame w_y, w_z stale (fpv_asymmetry.py:876) keeps its meaning.  Every operation is +, * on float32,
so a -fmad=false CUDA build reproduces the torch result bit for bit.
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


def integrate(pos, quat, linvel, w_b, force_b, torque_b, dt, substeps):
    """One simulate(dt) call.  All tensors (N,3|4) float32; w_b is the body-frame angular velocity.
    Returns the new (pos, quat, linvel, w_b)."""
    inertia = torch.tensor(INERTIA, dtype=torch.float32)
    inv_inertia = torch.tensor([1.0 / i for i in INERTIA], dtype=torch.float32)
    inv_mass = 1.0 / MASS
    h = dt / substeps
    hh = 0.5 * h
    grav = torch.tensor([0.0, 0.0, GRAVITY_Z], dtype=torch.float32)
    acc = qrot(quat, force_b) * inv_mass + grav
    for _ in range(substeps):
        linvel = linvel + h * acc
        w_b = w_b + h * ((torque_b - cross3(w_b, inertia * w_b)) * inv_inertia)
        pos = pos + h * linvel
        w2 = (w_b[:, 0:1] * w_b[:, 0:1] + w_b[:, 1:2] * w_b[:, 1:2]) + w_b[:, 2:3] * w_b[:, 2:3]
        t = w2 * (hh * hh)                                     # theta^2
        k = hh * (1.0 + t * (-1.0 / 6.0 + t * (1.0 / 120.0)))  # sin(theta)/|w|
        c = 1.0 + t * (-0.5 + t * (1.0 / 24.0))                # cos(theta)
        quat = qmul(quat, torch.cat((w_b * k, c), dim=1))
        n2 = ((quat[:, 0:1] * quat[:, 0:1] + quat[:, 1:2] * quat[:, 1:2]) + quat[:, 2:3] * quat[:, 2:3]) + quat[:, 3:4] * quat[:, 3:4]
        quat = quat * (1.5 - 0.5 * n2)
    return pos, quat, linvel, w_b

"""Actuation chain of the TACO MAV model: body-rate PID, control allocation, battery
sag, first-order rotor lag, aerodynamics, real->sim rotor remap.  Torch-float32 CPU
restatement, stateless functions over explicit state tensors.  TEST INFRASTRUCTURE.

Reference directory: IsaacGymEnvs/isaacgymenvs/tasks/control/ (CTRL/).  Golden vectors
from the reference modules: tests/golden/dynamics.npz (oracle/make_golden.py).
"""
import math

import torch

# --- body-rate PID gains, CTRL/angvel_control.py:17-60 ------------------------------
PID_KP = (27.5, 50.0, 200.0)
PID_KD = 0.5
PID_ERR_MAX = 400.0
PID_D_MAX = 150.0
PID_OUT_GAIN = 0.4

# --- battery constants, CTRL/battery_dynamics.py:19-30 ------------------------------
BAT_A = (4.35, -0.1102178, 0.0103368, -4.3778e-4)
BAT_B = (0.0015778, -7.7608e-5, 0.0069498)
BAT_R_MIN = 4.5
BAT_K = 0.00104846
BAT_TAU_RC = 3.3
BAT_EFF = 0.75
BAT_CELLS = 6
BAT_CAP = 1500.0

# --- rotor / aero nominal parameters, CTRL/thrust_dynamics.py:46-47,156-167 ----------
OMEGA_POLY = (0.0, 12.9466, 0.1872, -5.1220, 0.5906)
ROTOR_SAMPLE_TIME = 0.001
AERO_KF, AERO_KT = 1.13e-05, 0.05
AERO_DX, AERO_DY = -0.386, -0.53
AERO_KTH = 0.009


def rate_pid(setpoint, omega_body, prev_err, dt):
    """CTRL/angvel_control.py:67-88 (compute).  ki = 0 and Kf = 0 in the reference
    (:22-25,:32-35) so the integral and feed-forward terms contribute exactly zero;
    the dead integral state (quirk: accumulated but multiplied by 0) is not carried.
    Returns (torque_cmd (N,3), new_prev_err)."""
    kp = torch.tensor(PID_KP, dtype=torch.float32)
    err = torch.clip(setpoint - omega_body, -PID_ERR_MAX, PID_ERR_MAX)
    prev = torch.where(prev_err == 0, err, prev_err)
    p_term = kp * err
    d_term = torch.clip(PID_KD * ((err - prev) / dt), -PID_D_MAX, PID_D_MAX)
    # reference sums P + I + D + FF with I = FF = 0.0
    out = PID_OUT_GAIN * (p_term + 0.0 + d_term + 0.0)
    return out, err


def allocate(u):
    """CTRL/fpv_dynamics.py:35-46 (control_allocator).  u = (collective, roll, pitch, yaw)
    command; returns per-rotor throttle in [100, 1000]."""
    u0, u1, u2 = u[:, 0], u[:, 1], u[:, 2]
    u3 = torch.clip(u[:, 3], -u0 / 2, u0 / 2)
    # rows of the mixing matrix CTRL/fpv_dynamics.py:28-33 ; matmul accumulates left->right
    f = torch.stack((u0 - u1 + u2 - u3, u0 - u1 - u2 + u3, u0 + u1 - u2 - u3, u0 + u1 + u2 + u3), dim=1)
    f = f - torch.clamp((f - 1000).max(dim=1, keepdim=True)[0], 0)
    return torch.clip(f, 100, 1000)


def mech_power(rotor_speed):
    """fpv_asymmetry.py:614."""
    return torch.sum(400 * (rotor_speed * 2 * math.pi / 4500) ** 3, dim=1, keepdim=True)


def battery_step(p_mech, u1, e_c, t, dt, enabled=True):
    """CTRL/battery_dynamics.py:47-75 (sim_process).  State (u1, e_c, t) each (N,1).
    Returns (voltage (N,1), u1, e_c, t)."""
    if not enabled:
        return torch.full_like(u1, BAT_A[0]) * BAT_CELLS, u1, e_c, t
    t = t + dt
    p_c = p_mech / BAT_EFF / (BAT_CELLS * BAT_CAP)
    e_c = e_c + p_c * dt
    p_avg = e_c / t
    r0 = BAT_B[0] + BAT_B[1] * p_avg + BAT_B[2] * BAT_CAP
    r0 = torch.where(r0 > BAT_R_MIN, r0, torch.full_like(r0, BAT_R_MIN))
    a0 = torch.full_like(u1, BAT_A[0])
    u0 = a0 + BAT_A[1] * e_c + BAT_A[2] * e_c ** 2 + BAT_A[3] * e_c ** 3
    u1 = u1 + ((BAT_K * p_c - u1) / BAT_TAU_RC) * dt
    volt = 1 / 2 * (u0 - u1 + torch.sqrt((u0 - u1) ** 2 - 4 * r0 * p_c)) * BAT_CELLS
    return volt, u1, e_c, t


def rotor_step(volt, throttle, omega, poly, lag_gain):
    """CTRL/thrust_dynamics.py:52-66 + 80-86 (throttle_voltage2omega, omega_compute);
    omega_delay (:88-96) has depth 1 and is a pass-through.  ``lag_gain`` is
    sample_time / response_time (N,4), a per-episode constant (:84)."""
    x = throttle / 1000
    y = (volt - 23) / 3
    target = (poly[:, 0:1] * torch.ones_like(x) + poly[:, 1:2] * x + poly[:, 2:3] * y
              + poly[:, 3:4] * x ** 2 + poly[:, 4:5] * x * y) * 100
    return omega + lag_gain * (target - omega)


def aero_step(v_body, omega, aero):
    """CTRL/thrust_dynamics.py:173-199.  aero (N,5) = (k_f, k_tau, d_x, d_y, k_th).
    Returns rotor_force (N,4), rotor_torque (N,4), body_force (N,3)."""
    f = aero[:, 0:1] * omega * omega
    tq = aero[:, 1:2] * f
    vxy = torch.norm(v_body[:, :2], dim=1)
    body = torch.stack((aero[:, 2] * v_body[:, 0], aero[:, 3] * v_body[:, 1], aero[:, 4] * vxy * vxy), dim=1)
    return f, tq, body


def real_to_sim(f, tq):
    """CTRL/fpv_dynamics.py:48-56: sim rotor (0,1,2,3) <- real rotor (2,3,0,1); reaction
    torque negated on sim rotors 0 and 2."""
    fs = f[:, [2, 3, 0, 1]]
    ts = tq[:, [2, 3, 0, 1]].clone()
    ts[:, 0] = -ts[:, 0]
    ts[:, 2] = -ts[:, 2]
    return fs, ts

"""Pin the oracle's leaf functions to vectors produced by the reference's own torch
modules (oracle/make_golden.py; generated in the build container from /root/reference).
CPU only.  Bit-exact wherever the oracle repeats the reference's op sequence; the few
spots where the reference goes through BLAS (bmm in quat_rotate, matmul in the
allocator) are held to 2 ulp of the operand scale."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import leaf_math as lm
from oracle import dynamics as dyn
from oracle import rewards as rw


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def exact(a, b):
    assert a.shape == b.shape
    assert torch.equal(a, b), f"max abs diff {(a - b).abs().max().item():.3e}"


def test_leaf_math_vs_reference(golden_dir):
    g = _load(golden_dir, "leaf_math.npz")
    a, b, raw, v, eul = g["a"], g["b"], g["raw"], g["v"], g["eul"]
    exact(lm.qmul(a, b), g["quat_mul"])
    exact(lm.qconj(a), g["quat_conjugate"])
    # bmm-based dot in the reference: <= 2 ulp at |v| ~ 20
    torch.testing.assert_close(lm.qrot(a, v), g["quat_rotate"], rtol=0, atol=4e-6)
    torch.testing.assert_close(lm.qrot(lm.qconj(a), v), g["quat_rotate_conj"], rtol=0, atol=4e-6)
    exact(torch.stack(lm.euler_xyz(a), dim=1), g["euler"])
    exact(lm.roll_of(a), g["euler"][:, 0])
    exact(lm.quat_from_euler(eul[:, 0], eul[:, 1], eul[:, 2]), g["quat_from_euler"])
    exact(lm.rotmat9(raw), g["rotmat"])
    exact(lm.rotmat9(a), g["rotmat_unit"])
    exact(lm.quat_angle(a, b), g["quat_diff_rad"])
    exact(lm.rand_range(-math.pi, math.pi, g["u"]), g["rand_range_pi"])
    assert tuple(g["rand_float"].shape) == (4, 3)


def test_euler_gimbal_saturation(golden_dir):
    g = _load(golden_dir, "leaf_math.npz")
    # rows 0/1 were constructed with sin(pitch) ~ +-1 (asin branch within 1 ulp of saturation)
    assert abs(g["euler"][0, 1].item() - math.pi / 2) < 1e-3
    assert abs(g["euler"][1, 1].item() + math.pi / 2) < 1e-3
    sat = torch.tensor([[0.0, 0.70710678, 0.0, 0.70710679]], dtype=torch.float64).float() * 1.0000002
    assert abs(lm.euler_xyz(sat)[1].item() - math.pi / 2) < 1e-6      # |sinp| >= 1 -> copysign(pi/2)


def test_pid_vs_reference(golden_dir):
    g = _load(golden_dir, "dynamics.npz")
    prev = torch.zeros_like(g["pid_sp"][0])
    dt = float(np.float32(0.001))
    for i in range(g["pid_sp"].shape[0]):
        out, prev = dyn.rate_pid(g["pid_sp"][i], g["pid_w"][i], prev, dt)
        exact(out, g["pid_out"][i])
        exact(prev, g["pid_prev"][i])


def test_allocator_vs_reference(golden_dir):
    g = _load(golden_dir, "dynamics.npz")
    thr = dyn.allocate(g["alloc_u"].clone())
    torch.testing.assert_close(thr, g["alloc_thr"], rtol=0, atol=2.5e-4)   # matmul vs sequential sum, 2 ulp at 1000
    assert float(thr.min()) >= 100 and float(thr.max()) <= 1000


def test_battery_vs_reference(golden_dir):
    g = _load(golden_dir, "dynamics.npz")
    n = g["bat_ec0"].shape[0]
    u1, ec, t = torch.zeros(n, 1), g["bat_ec0"].clone(), torch.zeros(n, 1)
    dt = float(np.float32(0.001))
    for i in range(g["bat_pm"].shape[0]):
        volt, u1, ec, t = dyn.battery_step(g["bat_pm"][i], u1, ec, t, dt, True)
        ref = g["bat_volt"][i]
        both_nan = torch.isnan(volt) & torch.isnan(ref)
        assert torch.equal(torch.where(both_nan, torch.zeros_like(volt), volt), torch.where(both_nan, torch.zeros_like(ref), ref))
    exact(u1, g["bat_u1"]); exact(ec, g["bat_ec"]); exact(t, g["bat_t"])
    volt, *_ = dyn.battery_step(g["bat_pm"][1], u1, ec, t, dt, False)
    exact(volt, g["bat_off_volt"])
    assert abs(volt[0, 0].item() - 26.1) < 1e-5          # 4.35 V x 6 cells, battery_dynamics.py:19,28,75


def test_rotor_vs_reference(golden_dir):
    g = _load(golden_dir, "dynamics.npz")
    om = g["rot_om0"].clone()
    gain = dyn.ROTOR_SAMPLE_TIME / g["rot_tau"]
    for i in range(g["rot_volt"].shape[0]):
        om = dyn.rotor_step(g["rot_volt"][i], g["rot_thr"][i], om, g["rot_poly"], gain)
        exact(om, g["rot_om"][i])


def test_aero_remap_power_vs_reference(golden_dir):
    g = _load(golden_dir, "dynamics.npz")
    f, tq, bf = dyn.aero_step(g["aero_vb"], g["aero_om"], g["aero_par"])
    exact(f, g["aero_f"]); exact(tq, g["aero_tq"]); exact(bf, g["aero_bf"])
    assert float(g["aero_bt"].abs().max()) == 0.0        # body torque is identically zero in the reference
    fs, ts = dyn.real_to_sim(f, tq)
    exact(fs, g["remap_f"]); exact(ts, g["remap_tq"])
    exact(dyn.mech_power(g["aero_om"]), g["mech_power"])


def test_rewards_vs_reference(golden_dir):
    g = _load(golden_dir, "rewards.npz")
    r, x = rw.pos_reward(g["rel_body"], g["cpos"], g["cq"], g["tq"], g["prog"], 1000.0)
    exact(r, g["pos_rew"]); exact(x, g["pos_reset"])
    r, x = rw.rotate_reward(g["rel_world"].clone(), g["rel_vel"], g["cpos"], g["cq"], g["cmd_rot"], g["prog"], 1000.0)
    exact(r, g["rot_rew"]); exact(x, g["rot_reset"])
    r, x = rw.flip_reward(g["rel_body"], g["relq"], g["cpos"], g["cmd_flip"], g["prog"], 1000.0)
    exact(r, g["flip_rew"]); exact(x, g["flip_reset"])
    assert x.dtype == torch.int64
    # termination corner cases baked into the vectors: z<0.1, dist>10, progress>=max-1
    assert g["pos_reset"][:4].tolist() == [1, 1, 1, 1]
    assert g["pos_reset"][10:14].tolist() == [0, 1, 1, 0] or g["pos_reset"][11:13].tolist() == [1, 1]

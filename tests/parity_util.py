"""Shared helpers for the GPU parity tests: run RefFpvEnv (oracle, CPU) and FpvVecTask (CUDA,
through the C ABI) on the same Philox action stream and compare field by field."""
import numpy as np
import torch

from oracle import philox as px
from oracle.fpv_env import RefFpvEnv, TASK_FLIP, TASK_ROTATE

# export_state column layout (include/taco_b200.h TACO_STATE_WORDS, taco_env.cu export_state_kernel)
COLS = dict(pos=(0, 3), quat=(3, 7), linvel=(7, 10), angvel=(10, 13), tpos=(13, 16), tquat=(16, 20),
            roll_old=(20, 21), roll_cont=(21, 22), rotor=(22, 26), pid_prev=(26, 29), cmd=(29, 30),
            battery=(30, 33), ep_return=(33, 34), progress=(34, 35), delay_len=(35, 36), q_runs=(36, 37),
            overflow=(37, 38), reset=(38, 39), poly=(40, 45), aero=(45, 50), lag=(50, 54))
# characteristic magnitude of each float field: relative error is |a-b| / max(|b|, scale)
SCALE = dict(pos=1.0, quat=1.0, linvel=1.0, angvel=1.0, tpos=1.0, tquat=1.0, roll_old=1.0, roll_cont=1.0,
             rotor=100.0, pid_prev=1.0, cmd=1.0, battery=1.0, ep_return=1e-2, poly=1.0, aero=1e-2, lag=1e-2,
             obs=1.0, states=1.0, rew=1e-2)
INT_FIELDS = ("progress", "delay_len", "overflow", "reset")


def oracle_actions(env, t):
    """U(-1,1) actions from the shared Philox action stream."""
    return torch.from_numpy(px.u01(px.draw(env.seed, env.gid, t, 0, px.STREAM_ACTIONS)) * np.float32(2.0) - np.float32(1.0))


def oracle_state(env):
    """RefFpvEnv -> dict of numpy arrays named like COLS."""
    is_flip = (env.task == TASK_FLIP).numpy()
    is_rot = (env.task == TASK_ROTATE).numpy()
    cmd = np.where(is_flip, env.flip_radian.numpy(), np.where(is_rot, env.command[:, 1].numpy(), 0.0)).astype(np.float32)
    n = lambda t: t.numpy()
    return dict(pos=n(env.pos), quat=n(env.quat), linvel=n(env.linvel), angvel=n(env.angvel), tpos=n(env.tpos),
                tquat=n(env.tquat), roll_old=n(env.rpy_old[:, 0:1]), roll_cont=n(env.rpy_cont[:, 0:1]),
                rotor=n(env.rotor_speed), pid_prev=n(env.pid_prev), cmd=cmd[:, None],
                battery=np.concatenate([n(env.bat_u1), n(env.bat_ec), n(env.bat_t)], axis=1),
                ep_return=n(env.ep_return)[:, None], progress=n(env.progress_buf)[:, None],
                delay_len=n(env.delay_len)[:, None], overflow=n(env.overflow)[:, None].astype(np.int64),
                reset=n(env.reset_buf)[:, None], poly=n(env.poly), aero=n(env.aero), lag=n(env.lag_gain),
                is_flip=is_flip)


def rel_err(a, b, scale):
    """Relative error of a field: per env, max_i |a_i - b_i| / max(||b||_2, scale) where b is the
    env's whole vector for that field (3-vector, quaternion, 4 rotor speeds, one 26-value frame...),
    so that a small component of a large vector is measured against the vector, not against itself."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    norm = np.sqrt(np.sum(np.where(np.isfinite(b), b, 0.0) ** 2, axis=-1, keepdims=True)) if b.ndim > 1 else np.abs(b)
    err = np.abs(a - b) / np.maximum(norm, scale)
    err = np.where(both_nan, 0.0, err)
    return np.where(np.isnan(err), np.inf, err)


def elem_rel_err(a, b, floor):
    """Element-wise relative error max_i |a_i - b_i| / max(|b_i|, floor): the literal "relative error" of the north star, with a floor
    below which an element is measured absolutely (a component that is exactly zero in one implementation has no relative error)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    ok = np.isfinite(a) & np.isfinite(b)
    return float((np.abs(a - b)[ok] / np.maximum(np.abs(b)[ok], floor)).max()) if ok.any() else 0.0


def elementwise_errors(obs_gpu, rew_gpu, obs_ref, rew_ref, mask=None):
    """{obs, states, rew}: element-wise relative errors (floors 1e-2 of the field's unit scale)."""
    m = slice(None) if mask is None else np.asarray(mask)
    return dict(obs=elem_rel_err(obs_gpu["obs"].cpu().numpy()[m], obs_ref["obs"].numpy()[m], 1e-2),
                states=elem_rel_err(obs_gpu["states"].cpu().numpy()[m], obs_ref["states"].numpy()[m], 1e-2),
                rew=elem_rel_err(rew_gpu.cpu().numpy()[m], rew_ref.numpy()[m], 1e-4))


def compare(gpu_env, ref_env, obs_gpu, rew_gpu, reset_gpu, tout_gpu, obs_ref, rew_ref, reset_ref, tout_ref, mask=None):
    """Returns ({field: max relative error}, {int field: mismatch count}).  ``mask`` restricts the
    comparison to a subset of envs (e.g. finite, in-domain)."""
    st = gpu_env.export_state()
    ref = oracle_state(ref_env)
    N = st.shape[0]
    m = np.ones(N, dtype=bool) if mask is None else np.asarray(mask)
    errs, mism = {}, {}
    for name, (a, b) in COLS.items():
        if name in ("q_runs",):
            continue
        g = st[:, a:b]
        r = ref[name]
        if name in INT_FIELDS:
            mism[name] = int((g[m].astype(np.int64) != r[m].astype(np.int64)).sum())
            continue
        mm = m & ref["is_flip"] if name in ("roll_old", "roll_cont") else m
        if not gpu_env.cfg_has_dr and name in ("poly", "aero", "lag"):
            pass
        errs[name] = float(rel_err(g[mm], r[mm], SCALE[name]).max()) if mm.any() else 0.0
    errs["obs"] = float(rel_err(obs_gpu["obs"].cpu().numpy()[m], obs_ref["obs"].numpy()[m], SCALE["obs"]).max())
    errs["states"] = float(rel_err(obs_gpu["states"].cpu().numpy()[m], obs_ref["states"].numpy()[m], SCALE["states"]).max())
    errs["rew"] = float(rel_err(rew_gpu.cpu().numpy()[m], rew_ref.numpy()[m], SCALE["rew"]).max())
    mism["reset_buf"] = int((reset_gpu.cpu().numpy()[m] != reset_ref.numpy()[m]).sum())
    mism["time_outs"] = int((tout_gpu.cpu().numpy()[m] != tout_ref.numpy()[m]).sum())
    return errs, mism


def make_pair(cfg, seed=0x7AC0, strict_fp=True, debug_delay=True, env_offset=0, num_envs_global=None):
    import taco_b200
    gpu = taco_b200.FpvVecTask(cfg, "cuda:0", "cuda:0", -1, True, env_offset=env_offset, num_envs_global=num_envs_global,
                               seed=seed, strict_fp=strict_fp, debug_delay=debug_delay)
    gpu.cfg_has_dr = bool(cfg["random_rotordynamic_coe"] or cfg["random_rotor_response"] or cfg["random_aerodynamic_coe"])
    ref = RefFpvEnv(cfg, env_offset=env_offset, num_envs_global=num_envs_global, seed=seed)
    return gpu, ref


def run_lockstep(gpu, ref, steps, check_delay=True, on_step=None):
    """Step both envs on identical actions; returns per-step (errs, mism, delay_mismatch)."""
    out = []
    for t in range(steps):
        a = oracle_actions(ref, t)
        a_gpu = gpu.random_actions(t)
        assert torch.equal(a_gpu.cpu(), a), "Philox action stream differs between kernel and oracle"
        o_g, r_g, x_g, e_g = gpu.step(a_gpu)
        o_r, r_r, x_r, e_r = ref.step(a)
        finite = torch.isfinite(o_r["states"]).all(dim=2).all(dim=1).numpy() & ~ref.overflow.numpy()
        finite &= np.isfinite(oracle_state(ref)["rotor"]).all(axis=1)
        errs, mism = compare(gpu, ref, o_g, r_g, x_g, e_g["time_outs"], o_r, r_r, x_r, e_r["time_outs"], mask=finite)
        dmis = 0
        if check_delay:
            dd = gpu.debug_delay()                                             # (N, 10, 4)
            rows = torch.arange(ref.N)
            # oracle: the dense buffer was shifted by 10 after the reads, so re-derive from the log
            ref_dd = ref.last_delayed_actions.numpy()
            dmis = int((dd[finite] != ref_dd[finite]).sum())
        if t == 0:
            ref.first_step_elementwise = elementwise_errors(o_g, r_g, o_r, r_r, mask=finite)
        out.append((errs, mism, dmis, int(finite.sum())))
        if on_step:
            on_step(t, errs, mism, dmis)
    return out

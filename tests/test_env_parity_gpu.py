"""Parity of the fused CUDA step (through the C ABI / FpvVecTask) against the oracle (RefFpvEnv) on
identical Philox draws.  GPU only.

Bars (BASELINE.json north_star), for the default strict (-fmad=false) build -- the one bench.py times:
  * delay-buffer reads, reset masks, episode counters, time-outs: bit-exact, every step;
  * single-step state / obs / reward: <= 1e-5 relative FP32, both as max_i |a_i-b_i| / max(||b||_2, scale) per env and field
    (parity_util.rel_err / SCALE, every exported state field) and element-wise for obs / states / reward
    (parity_util.elem_rel_err).  Measured: <= 5e-7 (pos, quat, velocities, PID state are
    bit-identical after the first step; the rest is 1-ulp libm noise: atan2/asin, and torch's AVX-512 sqrt, which
    is not correctly rounded for ~0.7 % of inputs while CUDA's sqrt.rn is);
  * 50-step horizon: <= 5e-4 for every task (measured <= 3e-5).  The flip dynamics are chaotic, but the strict
    build performs the same float32 operations in the same order as the oracle, so the trajectories only separate
    through the 1-ulp sqrt differences above.
The FMA-contracted build (strict_fp=False) rounds differently by construction: it is held to 2e-5 on the first
step, and its long-horizon divergence (ill-conditioned battery sag near full throttle, battery_dynamics.py:68)
is reported by tools/parity_report.py rather than asserted.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SINGLE_STEP_TOL_STRICT = 1.0e-5
SINGLE_STEP_TOL_FAST = 2.0e-5
H50_TOL = 5.0e-4


def _pair(task, n=4096, strict=True, **kw):
    import parity_util as pu
    from taco_b200 import make_cfg
    cfg = make_cfg(task, n, **kw)
    return pu, *pu.make_pair(cfg, strict_fp=strict)


def _assert_ints(mism, dmis, where):
    assert all(v == 0 for v in mism.values()), f"{where}: integer/mask mismatch {mism}"
    assert dmis == 0, f"{where}: {dmis} delayed-action words differ"


@pytest.mark.parametrize("task", ["flip", "pos", "rotate", "mix"])
def test_single_step_and_50_step_horizon_strict(task):
    pu, gpu, ref = _pair(task)
    res = pu.run_lockstep(gpu, ref, 50)
    errs1, mism1, dmis1, nfin1 = res[0]
    assert nfin1 == 4096
    _assert_ints(mism1, dmis1, f"{task} step 1")
    assert max(errs1.values()) <= SINGLE_STEP_TOL_STRICT, f"{task} single-step: {errs1}"
    # the same bar ELEMENT-WISE (|a_i - b_i| / max(|b_i|, floor), floors 1e-2 for the O(1) frame values and 1e-4 for the reward):
    # measured <= 3.0e-6 / 4.4e-7 (profiles/parity_r02k.log)
    assert max(ref.first_step_elementwise.values()) <= SINGLE_STEP_TOL_STRICT, f"{task} single-step, element-wise: {ref.first_step_elementwise}"
    # second step: first step with non-zero wrench (the first step after a reset applies zero force, quirk 1)
    errs2 = res[1][0]
    assert max(errs2.values()) <= SINGLE_STEP_TOL_STRICT, f"{task} step 2: {errs2}"
    for t, (errs, mism, dmis, nfin) in enumerate(res):
        _assert_ints(mism, dmis, f"{task} step {t + 1}")
    errs50 = res[-1][0]
    assert max(errs50.values()) <= H50_TOL, f"{task} 50-step: {errs50}"
    # rollout statistics: integer counts exact, sums to float32 accumulation accuracy
    g = gpu.stats().cpu().numpy()
    s = ref.stats
    assert g[1] == s["n_done"] and g[2] == s["n_timeout"] and g[4] == s["sum_ep_len"] and g[7] == s["n_steps"]
    assert g[5] == s["n_nonfinite"] and g[6] == s["n_delay_overflow"]
    assert abs(g[0] - s["sum_reward"]) <= 1e-3 * abs(s["sum_reward"])
    gpu.close()


def test_single_step_fast_build_flip():
    pu, gpu, ref = _pair("flip", strict=False)
    res = pu.run_lockstep(gpu, ref, 1)
    for t, (errs, mism, dmis, nfin) in enumerate(res):
        _assert_ints(mism, dmis, f"flip fast step {t + 1}")
    assert max(res[0][0].values()) <= SINGLE_STEP_TOL_FAST, res[0][0]
    gpu.close()


def test_domain_randomisation_noise_random_delay_deploy():
    """BASELINE config 5 switches: per-env rotor response / polynomial / aero DR, random delay and deploy
    time, random voltage, plus observation noise and rotor noise."""
    pu, gpu, ref = _pair("mix", domain_randomization=True, observation_noise=True, rotor_noise=True)
    res = pu.run_lockstep(gpu, ref, 30)
    for t, (errs, mism, dmis, nfin) in enumerate(res):
        _assert_ints(mism, dmis, f"mix+DR step {t + 1}")
    assert max(res[0][0].values()) <= SINGLE_STEP_TOL_STRICT, res[0][0]
    st = gpu.export_state()
    lens = st[:, 35]
    assert lens.min() >= 0 and len(np.unique(lens)) > 3          # random delay + deploy -> spread of queue lengths
    gpu.close()


def test_history_is_not_cleared_on_reset_and_frames_shift():
    """Quirk 2 (fpv_asymmetry.py:392,413): states history keeps pre-reset frames; frame f of step t is
    frame f-1 of step t+1, bit for bit."""
    pu, gpu, ref = _pair("flip", n=2048)
    prev = None
    for t in range(12):
        a = gpu.random_actions(t)
        o, r, x, e = gpu.step(a)
        st = o["states"].clone()
        if prev is not None:
            assert torch.equal(st[:, :-1], prev[:, 1:])
        assert torch.equal(o["obs"][:, -1], st[:, -1])             # no observation noise: obs == newest state frame
        prev = st
    gpu.close()


def test_timeouts_and_episode_length():
    """progress >= max_len-1 => reset and time_out (vec_task_asymmetry.py:323); counters restart at 0."""
    import parity_util as pu
    from taco_b200 import make_cfg
    cfg = make_cfg("pos", 512)
    cfg["env"]["maxEpisodeLength"] = 6
    gpu, ref = pu.make_pair(cfg)
    res = pu.run_lockstep(gpu, ref, 14)
    for t, (errs, mism, dmis, nfin) in enumerate(res):
        _assert_ints(mism, dmis, f"timeout step {t + 1}")
    g = gpu.stats().cpu().numpy()
    assert g[2] == ref.stats["n_timeout"] and g[2] > 0
    gpu.close()


def test_command_redraw_at_progress_500():
    """reset_command_condition (fpv_asymmetry.py:587-603): flip_radian += 2*pi*k at progress == 500."""
    import parity_util as pu
    from taco_b200 import make_cfg
    cfg = make_cfg("flip", 256)
    gpu, ref = pu.make_pair(cfg, debug_delay=False)
    # jump both envs to progress 499 after the first (resetting) step, then step across 500
    a = pu.oracle_actions(ref, 0)
    gpu.step(gpu.random_actions(0)); ref.step(a)
    st = gpu.export_state()
    st[:, 34] = 499; st[:, 38] = 0
    gpu.import_state(st)
    ref.progress_buf[:] = 499; ref.reset_buf[:] = 0
    before = st[:, 29].copy()
    for t in range(1, 4):
        a = pu.oracle_actions(ref, t)
        gpu.step(gpu.random_actions(t)); ref.step(a)
    after = gpu.export_state()[:, 29]
    alive = (ref.progress_buf.numpy() == 502)
    assert alive.sum() > 10
    assert np.array_equal(after[alive], ref.flip_radian.numpy()[alive])
    k = np.round((after[alive] - before[alive]) / (2 * np.pi))
    assert set(np.unique(k)).issubset({-3, -2, -1, 0, 1, 2, 3}) and len(np.unique(k)) >= 4
    gpu.close()


def test_shard_invariance_single_gpu():
    """Env g gives the same trajectory whether it is simulated in a 4096-env handle or in a 1024-env shard
    at env_offset 2048 (Philox keyed by global env id; mix groups from global thirds)."""
    import parity_util as pu
    import taco_b200
    from taco_b200 import make_cfg
    full = taco_b200.FpvVecTask(make_cfg("mix", 4096, domain_randomization=True), "cuda:0", "cuda:0", -1, True, seed=7)
    shard = taco_b200.FpvVecTask(make_cfg("mix", 1024, domain_randomization=True), "cuda:0", "cuda:0", -1, True,
                                 env_offset=2048, num_envs_global=4096, seed=7)
    for t in range(20):
        af = full.random_actions(t)
        o_f, r_f, x_f, _ = full.step(af)
        o_s, r_s, x_s, _ = shard.step(af[2048:3072].contiguous())
        assert torch.equal(o_f["states"][2048:3072], o_s["states"])
        assert torch.equal(r_f[2048:3072], r_s) and torch.equal(x_f[2048:3072], x_s)
    assert np.array_equal(full.export_state()[2048:3072, :34], shard.export_state()[:, :34])
    full.close(); shard.close()


@pytest.mark.parametrize("task,dr,n,steps", [("flip", False, 24_001, 40), ("mix", True, 24_001, 40), ("flip", False, 2_097_152, 12), ("mix", True, 2_097_152, 8)])
def test_large_grid_launch_matches_small_grid_shards(task, dr, n, steps):
    """The step kernel has two launch shapes: grids of at most one CTA per SM (<= 148 x 128 envs: the 4096-env scale every oracle
    comparison in this file runs at) add four copy warps per CTA that move the kept state history while the step computes; larger
    grids -- the benchmarked shape -- copy after the step.  One 24 001-env handle (188 CTAs, ragged tail) against 4096-env shards
    of the same global envs: states, observations, rewards, masks and the exported env state must be bit-identical, step after
    step, so the oracle parity of the small shape carries over to the large one."""
    import taco_b200
    from taco_b200 import make_cfg
    # (the third case is bench.py's own workload: BASELINE configs[1] at 2 Mi envs per GPU)
    full = taco_b200.FpvVecTask(make_cfg(task, n, domain_randomization=dr), "cuda:0", "cuda:0", -1, True, seed=11)
    offs = [0, (n // 2 // 128) * 128 + 128, n - 4096]
    shards = [taco_b200.FpvVecTask(make_cfg(task, 4096, domain_randomization=dr), "cuda:0", "cuda:0", -1, True,
                                   env_offset=o, num_envs_global=n, seed=11) for o in offs]
    for t in range(steps):
        af = full.random_actions(t)
        o_f, r_f, x_f, e_f = full.step(af)
        for o, sh in zip(offs, shards):
            o_s, r_s, x_s, e_s = sh.step(af[o:o + 4096].contiguous())
            assert torch.equal(o_f["states"][o:o + 4096], o_s["states"]), (t, o)
            assert torch.equal(o_f["obs"][o:o + 4096], o_s["obs"]), (t, o)
            assert torch.equal(r_f[o:o + 4096], r_s) and torch.equal(x_f[o:o + 4096], x_s)
            assert torch.equal(e_f["time_outs"][o:o + 4096], e_s["time_outs"])
    st = full.export_state()
    for o, sh in zip(offs, shards):
        assert np.array_equal(st[o:o + 4096, :34], sh.export_state()[:, :34])
        sh.close()
    full.close()


@pytest.mark.parametrize("mode", ["mapped", "copy", "pageable"])
@pytest.mark.parametrize("n", [3000, 200_003])     # one chunk / three pipelined chunks; neither a multiple of 128 (tail block)
def test_host_buffer_entry_point_matches_device_path(n, mode, monkeypatch):
    """taco_env_step_host: 'mapped' = the kernel reads / writes the pinned host buffers itself (default), 'copy' = the
    chunked copy pipeline (TACO_HOST_MODE=copy), 'pageable' = unpinned buffers (falls back to the copy pipeline)."""
    import taco_b200
    from taco_b200 import make_cfg
    monkeypatch.setenv("TACO_HOST_MODE", "copy" if mode == "copy" else "mapped")
    pin = (lambda t: t) if mode == "pageable" else (lambda t: t.pin_memory())
    a_env = taco_b200.FpvVecTask(make_cfg("flip", n), "cuda:0", "cuda:0", -1, True, seed=3)
    b_env = taco_b200.FpvVecTask(make_cfg("flip", n), "cuda:0", "cuda:0", -1, True, seed=3)
    h_rew = pin(torch.empty(n)); h_reset = pin(torch.empty(n, dtype=torch.int64))
    h_tout = pin(torch.empty(n, dtype=torch.uint8))
    for t in range(5):
        act = a_env.random_actions(t)
        o, r, x, e = a_env.step(act)
        h_rew.fill_(-7.0); h_reset.fill_(-7); h_tout.fill_(77)
        b_env.step_host(pin(act.cpu()), h_rew, h_reset, h_tout)
        assert torch.equal(r.cpu(), h_rew) and torch.equal(x.cpu(), h_reset)
        assert torch.equal(e["time_outs"].cpu().to(torch.uint8), h_tout)
        assert torch.equal(o["states"], b_env.states_buf)
        assert torch.equal(r, b_env.rew_buf) and torch.equal(x, b_env.reset_buf)
    assert torch.equal(a_env.stats(), b_env.stats())
    a_env.close(); b_env.close()


def test_host_buffer_compact_result_format():
    """taco_env_step_host_compact: rew f32 + one flag byte per env (bit 0 reset, bit 1 time-out) equals the device-path results;
    pageable buffers are refused (mapped mode only)."""
    import taco_b200
    from taco_b200 import make_cfg
    n = 4097
    cfg = make_cfg("pos", n)
    cfg["env"]["maxEpisodeLength"] = 5                       # time-outs inside the run
    a_env = taco_b200.FpvVecTask(cfg, "cuda:0", "cuda:0", -1, True, seed=3)
    b_env = taco_b200.FpvVecTask(cfg, "cuda:0", "cuda:0", -1, True, seed=3)
    h_rew, h_flags = torch.empty(n).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    seen = 0
    for t in range(8):
        act = a_env.random_actions(t)
        o, r, x, e = a_env.step(act)
        h_rew.fill_(-7.0); h_flags.fill_(77)
        b_env.step_host_compact(act.cpu().pin_memory(), h_rew, h_flags)
        assert torch.equal(r.cpu(), h_rew)
        assert torch.equal((x.cpu() != 0).to(torch.uint8) | (e["time_outs"].cpu().to(torch.uint8) << 1), h_flags)
        assert torch.equal(o["states"], b_env.states_buf) and torch.equal(x, b_env.reset_buf)
        seen |= int(h_flags.max())
    assert seen == 3
    with pytest.raises(RuntimeError):
        b_env.step_host_compact(torch.zeros(n, 4), h_rew, h_flags)
    a_env.close(); b_env.close()


def test_host_buffer_entry_point_without_result_buffers():
    """Every result pointer of taco_env_step_host may be NULL (actions only): the step still runs and the device buffers hold the
    results -- in mapped and in copy mode."""
    import os
    import taco_b200
    from taco_b200 import make_cfg
    n = 5000
    a_env = taco_b200.FpvVecTask(make_cfg("mix", n, domain_randomization=True), "cuda:0", "cuda:0", -1, True, seed=9)
    b_env = taco_b200.FpvVecTask(make_cfg("mix", n, domain_randomization=True), "cuda:0", "cuda:0", -1, True, seed=9)
    h_rew = torch.empty(n).pin_memory()
    try:
        for t in range(6):
            act = a_env.random_actions(t)
            o, r, x, e = a_env.step(act)
            os.environ["TACO_HOST_MODE"] = "copy" if t % 2 else "mapped"
            if t < 3:
                b_env.step_host(act.cpu().pin_memory())
            else:
                b_env.step_host(act.cpu().pin_memory(), h_rew)                     # one result buffer only
                assert torch.equal(r.cpu(), h_rew)
            assert torch.equal(r, b_env.rew_buf) and torch.equal(x, b_env.reset_buf) and torch.equal(o["states"], b_env.states_buf)
    finally:
        os.environ.pop("TACO_HOST_MODE", None)
    a_env.close(); b_env.close()


def test_api_errors():
    import taco_b200
    from taco_b200 import make_cfg
    env = taco_b200.FpvFlip(make_cfg("flip", 64))
    with pytest.raises(ValueError):
        env.step(torch.zeros(63, 4, device="cuda"))
    with pytest.raises(TypeError):
        env.step(np.zeros((64, 4), dtype=np.float32))
    obs = env.reset()
    assert obs["obs"].shape == (64, 1, 26) and obs["states"].shape == (64, 5, 26) and float(obs["states"].abs().max()) == 0.0
    assert env.reset_buf.dtype == torch.int64 and int(env.reset_buf.sum()) == 64            # vec_task_asymmetry.py:246-247
    o, r, x, e = env.step(torch.zeros(64, 4, device="cuda").t().contiguous().t())           # non-contiguous input is copied
    assert e["time_outs"].dtype == torch.bool and r.dtype == torch.float32
    env.difficulty = 0.5
    assert env.difficulty == 0.5
    bad = make_cfg("flip", 64); bad["env"]["controlFrequencyInv"] = 7
    with pytest.raises(RuntimeError, match="control_freq_inv"):
        taco_b200.FpvFlip(bad)
    env.close()


@pytest.mark.parametrize("task,n,kw", [
    ("flip", 1, {}),                                                    # a single env
    ("mix", 127, {}),                                                   # ragged: below one CTA; mix thirds of 127
    ("rotate", 129, {}),                                                # ragged: one env into the second CTA
    ("pos", 1000, {"env.lenObservations": 5, "env.lenStates": 1}),      # history shapes other than the README's (1, 5)
    ("flip", 640, {"env.lenObservations": 3, "env.lenStates": 7}),
    ("flip", 512, {"delay_time": 0}),                                   # no delay: the run appended this step is read at once
    ("mix", 512, {"delay_time": 90}),                                   # the largest delay the reference can hold (90 + 10 slots)
    ("flip", 512, {"sim.substeps": 1}),
    ("rotate", 512, {"sim.substeps": 3}),                               # run-time sub-step count (not an unrolled variant)
    ("flip", 512, {"battery_consumption": False, "rotor_response": False}),
    ("mix", 512, {"random_copter_pos": False, "random_copter_quat": False, "random_copter_vel": False, "random_target_pos": False,
                  "random_target_yaw": False, "random_voltage": False, "random_rotor_speed": False, "random_command": False}),
    ("pos", 512, {"env.clipActions": 0.5}),
])
def test_edge_shapes_and_switches(task, n, kw):
    """Ragged / minimal env counts, other history lengths, delay extremes, sub-step counts and every reset switch off:
    integer fields bit-exact on every step, floats within the single-step bar on steps 1-2 and the horizon bar after 12."""
    pu, gpu, ref = _pair(task, n=n, **kw)
    res = pu.run_lockstep(gpu, ref, 12)
    for t, (errs, mism, dmis, nfin) in enumerate(res):
        _assert_ints(mism, dmis, f"{task} n={n} {kw} step {t + 1}")
        assert nfin == n
    assert max(res[0][0].values()) <= SINGLE_STEP_TOL_STRICT, res[0][0]
    assert max(res[1][0].values()) <= SINGLE_STEP_TOL_STRICT, res[1][0]
    assert max(res[-1][0].values()) <= H50_TOL, res[-1][0]
    gpu.close()


def test_delay_beyond_the_reference_buffer_is_flagged():
    """delay_time = 100: the first write already crosses slot 100 (fpv_asymmetry.py:329 truncates it silently); every env
    is flagged and counted instead of compared."""
    pu, gpu, ref = _pair("flip", n=256, delay_time=100)
    a = gpu.random_actions(0)
    gpu.step(a); ref.step(pu.oracle_actions(ref, 0))
    st = gpu.export_state()
    assert (st[:, 37] == 1).all() and bool(ref.overflow.all())
    assert gpu.stats().cpu().numpy()[6] == 256
    gpu.close()


def _hover_actions(ref):
    """A small cascaded controller (position -> tilt -> body rates, altitude -> throttle) evaluated on the ORACLE's state: keeps part
    of the envs alive for whole episodes, so that the lock-step run below crosses progress == 500 and natural time-outs.  The actions
    are just inputs: both implementations receive the same tensor."""
    import math
    from oracle.leaf_math import euler_xyz, qconj, qrot
    roll, pitch, _ = euler_xyz(ref.quat)
    qc = qconj(ref.quat)
    e = qrot(qc, ref.tpos - ref.pos)                       # position error in the body frame
    v = qrot(qc, ref.linvel)
    ax = torch.clamp(1.2 * e[:, 0] - 1.8 * v[:, 0], -3.0, 3.0)
    ay = torch.clamp(1.2 * e[:, 1] - 1.8 * v[:, 1], -3.0, 3.0)
    rate_x = torch.clamp(6.0 * (-ay / 9.81 - roll), -8.0, 8.0)
    rate_y = torch.clamp(6.0 * (ax / 9.81 - pitch), -8.0, 8.0)
    thr = torch.clamp(-0.45 + 0.35 * (ref.tpos[:, 2] - ref.pos[:, 2]) - 0.25 * ref.linvel[:, 2], -1.0, 1.0)
    return torch.stack((thr, rate_x / 20.0, rate_y / 20.0, torch.zeros_like(thr)), dim=1).to(torch.float32).contiguous()


def test_1100_step_lockstep_crosses_progress_500_and_natural_timeouts():
    """maxEpisodeLength = 1000 (train_fpv_asymmetry_ppo.py:342), 1100 steps in lock-step with the oracle under a stabilising
    controller: envs live through the progress == 500 command re-draw (fpv_asymmetry.py:152,587-603) and reach the time-out at
    progress 999 (vec_task_asymmetry.py:323) without any state surgery.  Integer fields / masks / delayed actions exact at every step;
    floats stay within 2e-3 over whole 1000-step episodes (the controller is evaluated on the oracle's state, so the kernel runs
    these actions open-loop)."""
    import parity_util as pu
    from taco_b200 import make_cfg
    n = 96
    gpu, ref = pu.make_pair(make_cfg("mix", n, random_copter_quat=False, random_copter_vel=False))
    worst, n_to, crossed = 0.0, 0, 0
    for t in range(1100):
        a = _hover_actions(ref)
        o_g, r_g, x_g, e_g = gpu.step(a.cuda())
        crossed += int((ref.progress_buf == 500).sum())                    # evaluated by the step just taken (pre-physics)
        o_r, r_r, x_r, e_r = ref.step(a)
        fin = torch.isfinite(o_r["states"]).all(dim=2).all(dim=1).numpy() & ~ref.overflow.numpy()
        errs, mism = pu.compare(gpu, ref, o_g, r_g, x_g, e_g["time_outs"], o_r, r_r, x_r, e_r["time_outs"], mask=fin)
        assert all(v == 0 for v in mism.values()), (t, mism)
        assert int((gpu.debug_delay()[fin] != ref.last_delayed_actions.numpy()[fin]).sum()) == 0, t
        worst = max(worst, max(errs.values()))
        n_to += int(e_r["time_outs"].sum())
    assert crossed > 0 and n_to > 0, (crossed, n_to)
    assert worst <= 2e-3, worst
    print("1100-step lock-step: worst relative error", worst, "time-outs", n_to, "envs at progress 500:", crossed)
    gpu.close()

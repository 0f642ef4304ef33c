#!/usr/bin/env python
"""bench.py -- env-steps/s of the fused FPV step (flip task) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): env-steps/s, flip task, device-timed, max over ranks.  One "step" = one
VecTask.step over every env of the job (10 x 1 ms control sub-steps + observation history + reward
+ masked reset per env).  Envs shard independently over the GPUs (weak scaling: fixed envs per GPU,
global Philox env ids); the only collective is one NCCL all-reduce of the 8-double rollout
statistics vector at the end of the timed rollout.

Prints ONE JSON line on rank 0.  `value` = device-resident throughput through FpvVecTask.step;
`e2e` = the same step through the C ABI's host-buffer entry point (pinned-host actions H2D,
rew/reset/time_outs D2H, every step); `roofline` = HBM roofline of the step kernel using the
algorithmic bytes of SURVEY.md section 8(d); `cpu_baseline` = the oracle port (the reference's torch
modules' arithmetic + restated glue) timed on this host's cores on a bounded sample.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line: whatever native libraries print to file descriptor 1 (NCCL writes its version banner
# there when NCCL_DEBUG is set) is sent to stderr, and emit() writes the line to the real stdout.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

ALGO_BYTES_NO_DR = 1493      # 16 + 637 + 416 + 424 (len_obs=1, len_states=5, no per-env DR)
ALGO_BYTES_DR = 1549
TASK = "flip"
SEED = 0x7AC0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=2 * 1024 * 1024)
    ap.add_argument("--task", default=TASK, choices=["pos", "rotate", "flip", "mix"])
    ap.add_argument("--dr", action="store_true", help="per-env domain randomisation (BASELINE config 5)")
    ap.add_argument("--fast-fp", action="store_true", help="time the FMA-contracted kernel build instead of the default strict (-fmad=false) one")
    ap.add_argument("--cpu-envs", type=int, default=4096)
    ap.add_argument("--cpu-steps", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-small", action="store_true")
    ap.add_argument("--no-numa", action="store_true", help="multi-GPU runs: do not pin each rank to its GPU's NUMA node (comparison runs)")
    ap.add_argument("--no-actor", action="store_true", help="skip the BASELINE config 3 point (mix task + actor MLP in the rollout loop)")
    ap.add_argument("--actor-envs", type=int, default=262144)
    ap.add_argument("--actor-hidden", default="256,256,256", help="actor hidden sizes (not pinned by the reference: the YAML is missing; our stated default)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0, t1):
        sm, mx, pw, reasons = [], [], [], set()
        for ts, line in self.rows:
            if ts < t0 - 0.05 or ts > t1 + 0.15:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:   # timed region shorter than the sampling period: fall back to every sample we have
            for ts, line in self.rows:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except Exception:
                    pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle port)
def cpu_baseline(task, n_envs, steps, warm=3, dr=False):
    """The reference's CPU path for this hot path = its torch control/reward modules driven by the
    restated FpvBase glue + our rigid-body stand-in (oracle/).  Timed with all host threads."""
    import torch
    from oracle.fpv_env import RefFpvEnv
    from oracle import philox as px
    from taco_b200.config import make_cfg
    import numpy as np
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = make_cfg(task, n_envs, domain_randomization=dr)
    env = RefFpvEnv(cfg, seed=SEED)
    acts = [torch.from_numpy(px.u01(px.draw(SEED, env.gid, t, 0, px.STREAM_ACTIONS)) * np.float32(2) - np.float32(1)) for t in range(warm + steps)]
    for t in range(warm):
        env.step(acts[t])
    t0 = time.perf_counter()
    for t in range(warm, warm + steps):
        env.step(acts[t])
    dt = time.perf_counter() - t0
    return {"value": n_envs * steps / dt, "unit": "env-steps/s", "cores": threads, "kind": "port",
            "sample": f"{task} task, {n_envs} envs x {steps} steps after {warm} warm-up, oracle port (torch CPU float32, {threads} threads), {dt:.1f} s",
            "ms_per_step": dt / steps * 1e3}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference
    env class itself cannot run anywhere: IsaacGym/PhysX binaries are absent).  Rank 0 only."""
    if rank != 0:
        return
    steps = max(1, min(args.steps, 200))
    warm = max(1, min(args.warmup, 5))
    cb = cpu_baseline(args.task, args.cpu_envs, steps, warm, args.dr)
    line = {
        "impl": "reference", "metric": "env-steps/s", "value": cb["value"], "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.task} task, reference CPU path (oracle port), bounded sample {args.cpu_envs} envs/step",
                   "len_obs": 1, "len_states": 5, "substeps": 2},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ ours
def time_steps(env, actions, steps, torch, dist, world):
    """K consecutive steps bracketed by barrier + synchronize; CUDA events on the launching stream."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps):
        env.step(actions[t % len(actions)])
    stats = env.stats()
    if world > 1:
        dist.all_reduce(stats)                       # the one collective of the path: 8 doubles per rollout
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
        dist.barrier()
    return ms, stats


def _timed(torch, fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def actor_rollout_point(torch, taco_b200, dev, n, hidden, strict_fp, peaks, iters=50):
    """BASELINE config 3: mix task, n envs, spectral-normalised actor MLP inference in the rollout loop
    (obs -> actor.act -> clipped action -> env.step), random-init policy (orthogonal, gains sqrt2 / 0.01,
    nets_asymmetry.py:41-55), weights projected once (lipschitz 4, ppo_asymmetry.py:398-404)."""
    sizes = [26] + list(hidden) + [4]
    gen = torch.Generator().manual_seed(SEED)
    ws, bs = [], []
    for l in range(len(sizes) - 1):
        w = torch.empty(sizes[l + 1], sizes[l])
        torch.nn.init.orthogonal_(w, gain=(2 ** 0.5 if l + 2 < len(sizes) else 0.01), generator=gen)
        ws.append(w)
        bs.append(torch.zeros(sizes[l + 1]))
    actor = taco_b200.ActorMLP(26, list(hidden), 4, device=dev)
    actor.load(ws, bs, lipschitz_const=4.0)
    env = taco_b200.FpvVecTask(taco_b200.make_cfg("mix", n), dev, dev, -1, True, seed=SEED, strict_fp=strict_fp)
    acts = [env.random_actions(t) for t in range(4)]
    for t in range(5):
        env.step(acts[t % 4])
    obs = env.obs_buf.clone()
    mean = torch.empty(n, 4, device=dev)
    tc = actor.tensor_cores_available
    ms_fp32 = _timed(torch, lambda: actor.forward(obs, tensor_cores=False, out=mean), max(iters // 5, 5))
    ms_tc = _timed(torch, lambda: actor.forward(obs, tensor_cores=True, out=mean), iters) if tc else None
    k = [0]

    def env_only():
        env.step(acts[k[0] % 4]); k[0] += 1

    def loop():
        _, clipped, _, _ = actor.act(env.obs_buf, k[0], seed=SEED, tensor_cores=tc)
        env.step(clipped); k[0] += 1

    ms_env = _timed(torch, env_only, iters)
    ms_loop = _timed(torch, loop, iters)
    # ---- the critic of the reference's training command (README.md:60-66: LSTM encoder over the 5 x 26 state history + MLP) joins
    # the loop: agent.act = actor sample + critic value (nets_asymmetry.py:326-352), then env.step
    crit = None
    try:
        c_hid, c_mlp = 64, list(hidden)
        cs = [c_hid] + c_mlp + [1]
        lstm = []
        w_ih = torch.empty(4 * c_hid, 26); w_hh = torch.empty(4 * c_hid, c_hid)
        torch.nn.init.xavier_uniform_(w_ih, generator=gen); torch.nn.init.xavier_uniform_(w_hh, generator=gen)      # LSTMEncoder.para_init
        lstm.append((w_ih, w_hh, torch.zeros(4 * c_hid), torch.zeros(4 * c_hid)))
        cw, cb = [], []
        for l in range(len(cs) - 1):
            w = torch.empty(cs[l + 1], cs[l])
            torch.nn.init.orthogonal_(w, gain=(2 ** 0.5 if l + 2 < len(cs) else 0.01), generator=gen)
            cw.append(w); cb.append(torch.zeros(cs[l + 1]))
        critic = taco_b200.CriticLSTM(26, 5, c_hid, c_mlp, device=dev)
        critic.load(lstm, cw, cb)
        value = torch.empty(n, 1, device=dev)
        ctc = critic.tensor_cores_available
        ms_c_fp32 = _timed(torch, lambda: critic.forward(env.states_buf, tensor_cores=False, out=value), 3, warm=1)
        ms_c_tc = _timed(torch, lambda: critic.forward(env.states_buf, tensor_cores=True, out=value), iters) if ctc else None

        def loop_ac():
            _, clipped, _, _ = actor.act(env.obs_buf, k[0], seed=SEED, tensor_cores=tc)
            critic.forward(env.states_buf, tensor_cores=ctc, out=value)
            env.step(clipped); k[0] += 1

        ms_loop_ac = _timed(torch, loop_ac, iters)
        c_flops = 5 * 2.0 * (26 + c_hid) * 4 * c_hid + 2.0 * sum(cs[i] * cs[i + 1] for i in range(len(cs) - 1))
        crit = {"workload": f"critic = LSTM(26 -> {c_hid}) over the 5-frame state history + MLP {'x'.join(map(str, cs))} (sizes are OUR stated default: "
                            "the reference YAML is missing), agent.act = actor sample + critic value, then env.step, every step",
                "value": n / (ms_loop_ac * 1e-3), "unit": "env-steps/s", "ms_per_step": ms_loop_ac, "critic_fp32_ms": ms_c_fp32,
                "critic_tc_ms": ms_c_tc, "critic_flops_per_env": c_flops, "gpu_launches_per_step": 3}
        if ms_c_tc:
            ach = c_flops * n / (ms_c_tc * 1e-3) / 1e12
            peak = float(peaks.get("bf16_tflops", 1590.0))
            crit["critic_roofline"] = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                                       "kernel": "critic_tc_kernel", "speedup_vs_fp32_kernel": ms_c_fp32 / ms_c_tc}
        critic.close()
    except Exception as exc:      # the config-3 line must not be lost to the (newer) critic point
        crit = {"error": repr(exc)}
    flops = 2.0 * sum(sizes[i] * sizes[i + 1] for i in range(len(sizes) - 1))        # per env, SURVEY.md section 8(d)
    out = {"workload": f"mix task, {n} envs, actor {'x'.join(map(str, sizes))} (hidden sizes are OUR stated default: the reference YAML is missing), "
                       "random-init policy, spectral projection c=4 once per update, act -> clip -> step every step",
           "value": n / (ms_loop * 1e-3), "unit": "env-steps/s", "ms_per_step": ms_loop, "env_step_ms": ms_env,
           "actor_fp32_ms": ms_fp32, "actor_tc_ms": ms_tc, "actor_flops_per_env": flops, "gpu_launches_per_step": 2,
           "with_critic": crit}
    if ms_tc:
        ach = flops * n / (ms_tc * 1e-3) / 1e12
        peak = float(peaks.get("bf16_tflops", 1590.0))
        out["actor_roofline"] = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                                 "peak_source": "measured burst (MEASURED_PEAKS.json bf16_tflops)" if "bf16_tflops" in peaks else "fallback (B200_PROFILING.md)",
                                 "kernel": "actor_tc_kernel", "speedup_vs_fp32_kernel": ms_fp32 / ms_tc}
    env.close(); actor.close()
    return out


def small_rollout_point(torch, taco_b200, dev, task, hidden, strict_fp, n=4096, horizon=32, reps=5):
    """The rollout loop at the reference's own scale (README.md:41: 4096 envs): agent.act (actor sample + LSTM critic value) + env.step
    + zero-copy store + GAE, eager (collect_rollout) against one CUDA-graph replay per rollout (GraphedRollout, critic chain as a
    parallel branch).  Random-init networks of the bench's stated default sizes."""
    gen = torch.Generator().manual_seed(SEED)
    sizes = [26] + list(hidden) + [4]
    ws = [torch.randn(sizes[i + 1], sizes[i], generator=gen) / sizes[i] ** 0.5 for i in range(len(sizes) - 1)]
    bs = [torch.zeros(sizes[i + 1]) for i in range(len(sizes) - 1)]
    c_hid, cs = 64, [64] + list(hidden) + [1]
    lstm = [(torch.randn(4 * c_hid, 26, generator=gen) * 0.2, torch.randn(4 * c_hid, c_hid, generator=gen) * 0.15, torch.zeros(4 * c_hid), torch.zeros(4 * c_hid))]
    cw = [torch.randn(cs[i + 1], cs[i], generator=gen) / cs[i] ** 0.5 for i in range(len(cs) - 1)]
    cb = [torch.zeros(cs[i + 1]) for i in range(len(cs) - 1)]
    out = {"workload": f"{task}, {n} envs, horizon {horizon}: actor {'x'.join(map(str, sizes))} + critic LSTM 64 / MLP {'x'.join(map(str, cs))} in the loop, GAE at the end"}
    for mode in ("eager", "graph"):
        env = taco_b200.FpvVecTask(taco_b200.make_cfg(task, n), dev, dev, -1, True, seed=SEED, strict_fp=strict_fp)
        actor = taco_b200.ActorMLP(26, list(hidden), 4, device=dev); actor.load(ws, bs, lipschitz_const=-1.0)
        critic = taco_b200.CriticLSTM(26, 5, c_hid, list(hidden), device=dev); critic.load(lstm, cw, cb)
        buf = taco_b200.RolloutBuffer(n, 26, 1, 26, 5, 4, horizon, 1, 0.99, 0.95, dev)
        tc = actor.tensor_cores_available and critic.tensor_cores_available
        if mode == "eager":
            run = lambda: taco_b200.collect_rollout(env, actor, buf, critic, seed=SEED, tensor_cores=tc)
            run()
        else:
            gr = taco_b200.GraphedRollout(env, actor, buf, critic, seed=SEED, tensor_cores=tc)
            run = gr.run
        run()
        ms = _timed(torch, run, reps, warm=1)
        out[mode + "_ms_per_rollout"] = ms
        out[mode + "_env_steps_per_s"] = n * horizon / (ms * 1e-3)
        out[mode + "_us_per_step"] = ms / horizon * 1e3
        if mode == "graph":
            gr.close()
        env.close(); actor.close(); critic.close()
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    import torch
    import torch.distributed as dist
    import taco_b200
    from taco_b200 import dist as tdist
    numa_cpus = tdist.bind_to_gpu_numa(local_rank) if (world > 1 and not args.no_numa) else None    # pinned host buffers next to the rank's GPU
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback for the product path)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.envs_per_gpu
    cfg = taco_b200.make_cfg(args.task, n, domain_randomization=args.dr)
    dev = f"cuda:{local_rank}"
    env = taco_b200.FpvVecTask(cfg, dev, dev, -1, True, env_offset=rank * n, num_envs_global=world * n, seed=SEED, strict_fp=not args.fast_fp)
    n_act = 4
    actions = [env.random_actions(t) for t in range(n_act)]     # synthetic U(-1,1), resident in HBM before timing
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for t in range(max(args.warmup, 3)):
        env.step(actions[t % n_act])
    warm_stats = env.stats()
    if world > 1:
        dist.all_reduce(warm_stats)                  # NCCL communicator set-up happens here, outside the timed region
    t0 = sampler.mark()
    ms, stats = time_steps(env, actions, args.steps, torch, dist, world)
    t1 = sampler.mark()
    total_env_steps = float(world) * n * args.steps
    value = total_env_steps / (ms * 1e-3)
    # ---- end-to-end through the host-buffer C-ABI call
    e2e = None
    if not args.no_e2e:
        k_e2e = min(args.steps, 20)
        h_act = [actions[t % n_act].cpu().pin_memory() for t in range(n_act)]
        h_rew = torch.empty(n, dtype=torch.float32).pin_memory()
        h_reset = torch.empty(n, dtype=torch.int64).pin_memory()
        h_tout = torch.empty(n, dtype=torch.uint8).pin_memory()
        def time_host_steps(k):
            for t in range(2):
                env.step_host(h_act[t % n_act], h_rew, h_reset, h_tout)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for t in range(k):
                env.step_host(h_act[t % n_act], h_rew, h_reset, h_tout)
            e1.record()
            torch.cuda.synchronize()
            ms_ = e0.elapsed_time(e1)
            if world > 1:
                tm = torch.tensor([ms_], dtype=torch.float64, device="cuda")
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                ms_ = float(tm.item())
            return float(world) * n * k / (ms_ * 1e-3)
        os.environ["TACO_HOST_MODE"] = "copy"          # the chunked copy pipeline (what pageable buffers get), for comparison
        v_copy = time_host_steps(min(k_e2e, 10))
        os.environ["TACO_HOST_MODE"] = "mapped"        # default: the kernel reads / writes the pinned host buffers itself
        v_map = time_host_steps(k_e2e)
        e2e = {"value": v_map, "unit": "env-steps/s", "steps": k_e2e,
               "h2d_bytes_per_step": n * 16, "d2h_bytes_per_step": n * 13, "copy_pipeline_value": v_copy,
               "note": "per GPU bytes; taco_env_step_host with pinned host buffers, every step: one launch whose threads load their "
                       "action from host memory and post rew/reset/time_outs into host memory across PCIe, then a stream sync; "
                       "copy_pipeline_value = the same call with TACO_HOST_MODE=copy (chunked cudaMemcpyAsync pipeline)"}
    sampler.stop()
    clocks = sampler.summary(t0, t1) if rank == 0 else None
    # ---- small-N point of BASELINE config 2 (4096 envs, launch-latency bound)
    small = None
    if not args.no_small and world == 1:
        cfg_s = taco_b200.make_cfg(args.task, 4096)
        env_s = taco_b200.FpvVecTask(cfg_s, dev, dev, -1, True, seed=SEED, strict_fp=not args.fast_fp)
        acts_s = [env_s.random_actions(t) for t in range(n_act)]
        for t in range(10):
            env_s.step(acts_s[t % n_act])
        ms_s, _ = time_steps(env_s, acts_s, 200, torch, dist, 1)
        small = {"workload": f"{args.task}, 4096 envs (BASELINE config 2), L2-resident; 32 CTAs on 148 SMs = one warp per scheduler: bound by the latency of ~8.4k instructions per env-step, not by bandwidth", "value": 4096 * 200 / (ms_s * 1e-3),
                 "unit": "env-steps/s", "us_per_step": ms_s / 200 * 1e3}
        env_s.close()
        if not args.no_actor:
            try:
                small["rollout_loop"] = small_rollout_point(torch, taco_b200, dev, args.task, [int(x) for x in args.actor_hidden.split(",")], not args.fast_fp)
            except Exception as exc:
                small["rollout_loop"] = {"error": repr(exc)}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    actor_pt = None
    if not args.no_actor and world == 1:
        env.close()
        actor_pt = actor_rollout_point(torch, taco_b200, dev, args.actor_envs, [int(x) for x in args.actor_hidden.split(",")], not args.fast_fp, peaks)
    if rank == 0:
        if "hbm_gbs" in peaks:
            peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        algo = ALGO_BYTES_DR if args.dr else ALGO_BYTES_NO_DR
        kernel_ms = ms / args.steps
        achieved = algo * n / (kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_bytes_per_env.json")
        if os.path.exists(tp):
            try:
                traffic = float(json.load(open(tp)).get(args.task)) * n
            except Exception:
                traffic = None
        line = {
            "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.task} task fused env step (BASELINE configs[1] scaled to {n} envs/GPU so the working set exceeds L2), "
                                   f"len_obs=1, len_states=5, delay_time=20, rotor_response_time=0.017, substeps=2, "
                                   f"{'per-env DR on' if args.dr else 'no per-env DR'}, U(-1,1) Philox actions",
                       "envs_per_gpu": n, "global_envs": world * n, "parallelism": f"env-sharded x{world}", "rank0_numa_bound_cpus": (len(numa_cpus) if numa_cpus else None),
                       "l2": "inputs larger than L2 (working set ~%.1f GB/GPU)" % (n * 1900 / 1e9),
                       "fp_mode": "fast (FMA contraction)" if args.fast_fp else "strict (-fmad=false, the parity-tested build)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "fpv_step_kernel", "algorithmic_bytes_per_env_step": algo, "peak_source": peak_src,
                         "kernel_ms": kernel_ms},
            "gpu_launches": args.steps * world + 1 * world,
            "clocks": clocks,
            "rollout_stats": [float(x) for x in stats.cpu().tolist()],
        }
        if e2e is not None:
            line["e2e"] = e2e
        if small is not None:
            line["config2_4096_envs"] = small
        if actor_pt is not None:
            line["config3_mix_actor"] = actor_pt
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline(args.task, args.cpu_envs, args.cpu_steps, 3, args.dr).items() if k != "ms_per_step"}
        emit(line)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

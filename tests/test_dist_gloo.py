"""world_size-2 gloo test of the multi-GPU host logic (taco_b200/dist.py): sharding by global env id and the single
all-reduce of the rollout statistics vector.  Each rank steps an oracle shard (CPU); the reduced statistics must equal
those of the unsharded run, and per-env trajectories must be identical (shard invariance).  CPU only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from taco_b200 import dist as tdist
from taco_b200.config import make_cfg

N_PER_RANK, STEPS, SEED = 48, 4, 11


def _stats_vec(env):
    s = env.stats
    return torch.tensor([s["sum_reward"], s["n_done"], s["n_timeout"], s["sum_ep_return"], s["sum_ep_len"], s["n_nonfinite"],
                         s["n_delay_overflow"], s["n_steps"]], dtype=torch.float64)


def _run(env, steps):
    from parity_util import oracle_actions
    last = None
    for t in range(steps):
        last = env.step(oracle_actions(env, t))
    return last


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    from oracle.fpv_env import RefFpvEnv
    off, n_glob = tdist.shard(rank, world, N_PER_RANK)
    env = RefFpvEnv(make_cfg("mix", N_PER_RANK, domain_randomization=True), env_offset=off, num_envs_global=n_glob, seed=SEED)
    obs, rew, reset, extras = _run(env, STEPS)
    stats = tdist.allreduce_rollout_stats(_stats_vec(env))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), states=obs["states"].numpy(), rew=rew.numpy(), reset=reset.numpy(),
             stats=stats.numpy(), task=env.task.numpy())
    dist.destroy_process_group()


def test_two_rank_shards_match_single_process(tmp_path):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from oracle.fpv_env import RefFpvEnv
    full = RefFpvEnv(make_cfg("mix", 2 * N_PER_RANK, domain_randomization=True), seed=SEED)
    obs, rew, reset, _ = _run(full, STEPS)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    # torch-CPU vector kernels treat the tail of a tensor with scalar libm (sin/cos/atan2/sqrt), so the ORACLE is
    # shard-invariant only to 1 ulp; the CUDA kernel is bit-invariant (tests/test_env_parity_gpu.py::test_shard_invariance_single_gpu)
    st = np.concatenate([r0["states"], r1["states"]])
    assert np.abs(st - obs["states"].numpy()).max() <= 2e-5, np.abs(st - obs["states"].numpy()).max()
    assert np.allclose(np.concatenate([r0["rew"], r1["rew"]]), rew.numpy(), rtol=1e-4, atol=1e-8)
    assert np.array_equal(np.concatenate([r0["reset"], r1["reset"]]), reset.numpy())
    assert np.array_equal(np.concatenate([r0["task"], r1["task"]]), full.task.numpy())          # mix groups from global ids
    want = _stats_vec(full).numpy()
    for r in (r0, r1):
        assert np.allclose(r["stats"], want, rtol=1e-5, atol=1e-9)
        assert r["stats"][7] == 2 * N_PER_RANK * STEPS


def test_shard_and_groups_and_summary():
    assert tdist.shard(3, 8, 2097152) == (3 * 2097152, 16777216)
    with pytest.raises(ValueError):
        tdist.shard(8, 8, 4)
    assert tdist.mix_groups(4096) == [0, 1365, 2730, 4096] and tdist.mix_groups(16777216) == [0, 5592405, 11184810, 16777216]
    s = tdist.summarise(torch.tensor([10.0, 4, 1, 8.0, 400.0, 0, 0, 1000.0], dtype=torch.float64))
    assert s["mean_reward"] == 0.01 and s["mean_episode_return"] == 2.0 and s["mean_episode_length"] == 100.0
    with pytest.raises(ValueError):
        tdist.allreduce_rollout_stats(torch.zeros(8))
    assert tdist.allreduce_rollout_stats(torch.ones(8, dtype=torch.float64)).sum() == 8     # no process group: no-op


def test_bind_to_gpu_numa_is_harmless_without_nvml():
    """No GPU / no NVML here: the helper must return None and leave the affinity alone."""
    import os
    from taco_b200 import dist as tdist
    before = os.sched_getaffinity(0)
    cpus = tdist.bind_to_gpu_numa(0)
    assert cpus is None or cpus <= before
    if cpus is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)

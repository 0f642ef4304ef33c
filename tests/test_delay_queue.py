"""Property test: the run-length action queue reproduces the reference's dense delay buffer read for read, for every
delay in [0, 100] and every deploy-time sequence in {9,10,11}, as long as the reference itself stays inside its
delay_time_max (len + T <= 100); the overflow flag fires exactly when the reference truncates a write.  CPU only;
the CUDA implementation of the same algorithm is checked word for word in tests/test_env_parity_gpu.py."""
import numpy as np
from hypothesis import given, settings, strategies as st

from queue_model import DenseDelay, RunQueue


@settings(max_examples=300, deadline=None)
@given(delay=st.integers(0, 100), deploys=st.lists(st.integers(9, 11), min_size=1, max_size=120), seed=st.integers(0, 2**31 - 1), t0=st.integers(0, 2**32 - 1))
def test_run_queue_matches_dense_buffer(delay, deploys, seed, t0):
    rng = np.random.default_rng(seed)
    dense, rq = DenseDelay(delay), RunQueue(delay, t0)
    for t, T in enumerate(deploys):
        a = rng.uniform(-1, 1, 4).astype(np.float32)
        r_dense = dense.step(a, T)
        r_queue = rq.step(a, T)
        assert rq.overflow == dense.overflow
        if dense.overflow:
            return                       # outside delay_time_max: flagged, not compared
        assert rq.len == dense.len
        for k in range(10):
            assert np.array_equal(r_dense[k], r_queue[k]), (t, k)
        assert rq.n <= 13


def test_fixed_delay_20_reads_two_steps_late():
    dense, rq = DenseDelay(20), RunQueue(20)
    acts = [np.full(4, float(i + 1), dtype=np.float32) for i in range(6)]
    for i, a in enumerate(acts):
        rd, rq_ = dense.step(a, 10), rq.step(a, 10)
        want = 0.0 if i < 2 else float(i - 1)
        assert all(r[0] == want for r in rd) and all(r[0] == want for r in rq_)


def test_zero_delay_short_deploy_rereads_last_slot():
    # len = 0, T = 9: sub-step 9 reads slot min(len-1, 9) = 8, the same action (fpv_asymmetry.py:366)
    dense, rq = DenseDelay(0), RunQueue(0)
    a = np.arange(4, dtype=np.float32) + 1
    assert all(np.array_equal(x, a) for x in dense.step(a, 9)) and all(np.array_equal(x, a) for x in rq.step(a, 9))

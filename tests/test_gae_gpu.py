"""GAE / advantage-normalisation kernels (through RolloutBuffer and the C ABI) against the golden vectors produced by the
reference's own PPOReplayBuffer and against the oracle on larger seeded inputs.  GPU only.

Bars: ret and the un-normalised advantages are bit-exact (same float32 operations in the same order, -fmad=false); the
normalised advantages are within 2e-6 absolute of the reference's (adv - mean) / (std + 1e-8) (the kernel takes mean / std
from float64 moments, torch from float32 reductions)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "gae.npz"))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def _fill(buf, rew, done, value, time_outs=None):
    H, N = rew.shape[:2]
    z = lambda *s: torch.zeros(*s, device="cuda")
    for s in range(H):
        buf.store(z(N, buf.obs_len, buf.obs_dim), z(N, buf.states_len, buf.states_dim), z(N, 4), rew[s].cuda(), z(N), done[s].cuda(), value[s].cuda(),
                  z(N, 4), z(N, 4), time_outs=None if time_outs is None else time_outs[s].cuda())


@pytest.mark.parametrize("fused_bootstrap", [False, True])
def test_matches_reference_golden(golden_dir, fused_bootstrap):
    from taco_b200 import RolloutBuffer
    g = _golden(golden_dir)
    H, N = g["rew"].shape
    buf = RolloutBuffer(N, 26, 1, 26, 5, 4, H, 4, float(g["gamma"]), float(g["lam"]), "cuda:0")
    if fused_bootstrap:
        _fill(buf, g["rew"], g["done"], g["value"], g["time_outs"])          # raw reward + time_outs: bootstrap on the device
    else:
        _fill(buf, g["rew_aug"], g["done"], g["value"])                      # reward augmented by the caller, like the reference
    buf.compute_returns_and_advantage(g["last_value"].cuda())
    assert torch.equal(buf.ret_buf.cpu(), g["ret"])
    torch.testing.assert_close(buf.adv_buf.cpu(), g["adv_norm"], rtol=0, atol=2e-6)
    m = buf.moments.cpu()
    assert m[2].item() == H * N


@pytest.mark.parametrize("H,N", [(1, 1), (5, 130), (32, 4096 + 37), (20, 262144)])
def test_matches_oracle_on_seeded_inputs(H, N):
    from taco_b200 import RolloutBuffer
    from oracle import gae as og
    gen = torch.Generator().manual_seed(H * 1000 + N)
    rew = torch.rand(H, N, generator=gen) * 0.02
    value = torch.randn(H, N, 1, generator=gen) * 0.3 + 0.4
    done = (torch.rand(H, N, generator=gen) < 0.05).float()
    tout = (torch.rand(H, N, generator=gen) < 0.3)
    last = torch.randn(N, 1, generator=gen) * 0.3
    buf = RolloutBuffer(N, 26, 1, 26, 1, 4, H, 1, 0.99, 0.95, "cuda:0")
    _fill(buf, rew, done, value, tout)
    buf.compute_returns_and_advantage(last.cuda())
    aug = og.bootstrap_timeouts(rew, value.squeeze(-1), done, tout, 0.99)
    adv, ret = og.gae(aug.unsqueeze(-1), done.unsqueeze(-1), value, last, 0.99, 0.95)
    assert torch.equal(buf.ret_buf.cpu(), ret)
    if H * N > 1:
        want = og.normalize(adv)
        torch.testing.assert_close(buf.adv_buf.cpu(), want, rtol=0, atol=5e-6)
        torch.testing.assert_close(buf.moments.cpu(), og.moments(adv), rtol=1e-12, atol=1e-9)


def test_c_abi_rejects_bad_arguments():
    import ctypes as C
    from taco_b200 import _capi
    L = _capi.lib()
    assert L.taco_gae_advantages(0, 4, 4, None, None, None, None, None, 0.99, 0.95, None, None, None, None) == -1
    assert b"null" in L.taco_last_error()
    x = torch.zeros(16, device="cuda"); m = torch.zeros(3, dtype=torch.float64, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    assert L.taco_gae_advantages(0, 0, 4, p(x), p(x), None, p(x), p(x), 0.99, 0.95, p(x), p(x), p(m), None) == -1
    assert L.taco_gae_normalize(0, p(x), 0, p(m), None) == -1

"""Pin oracle/actor.py to the reference's own nets_asymmetry / ppo_asymmetry classes through
tests/golden/actor.npz (oracle/make_golden.py: actor()).  CPU only."""
import math
import os

import numpy as np
import torch

from oracle import actor as oa


def _load(golden_dir):
    z = np.load(os.path.join(golden_dir, "actor.npz"))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def _params(g, suffix=""):
    n = len([k for k in g if k.startswith("b")])
    return [g[f"w{i}{suffix}"] for i in range(n)], [g[f"b{i}"] for i in range(n)]


def test_mlp_forward_vs_reference(golden_dir):
    g = _load(golden_dir)
    w, b = _params(g)
    torch.testing.assert_close(oa.mlp_forward(g["obs"], w, b), g["mean"], rtol=0, atol=1e-6)   # same ops, BLAS blocking may differ
    torch.testing.assert_close(g["act_mean"], g["mean"], rtol=0, atol=1e-6)


def test_act_vs_reference(golden_dir):
    """Injecting the noise torch drew reproduces MultivariateNormal's sample and log_prob; std = exp(2 log_std)."""
    g = _load(golden_dir)
    action, clipped, logp = oa.act(g["act_mean"], g["log_std"], g["act_eps"])
    torch.testing.assert_close(action, g["act_action"], rtol=0, atol=2e-6)
    torch.testing.assert_close(logp, g["act_logp"], rtol=0, atol=2e-5)
    assert torch.equal(clipped, action.clamp(-1, 1))
    std = (g["act_action"] - g["act_mean"]).std(dim=0)
    expect = torch.exp(2 * g["log_std"])
    assert torch.all((std / expect - 1).abs() < 0.35)          # 96 samples: loose, but exp(log_std) would be off by 1.6x on column 0


def test_spectral_projection_vs_reference(golden_dir):
    g = _load(golden_dir)
    w, b = _params(g)
    c = float(g["lipschitz"])
    proj, sig = oa.spectral_normalize(w, c)
    torch.testing.assert_close(torch.tensor(sig), g["sigma_before"], rtol=1e-6, atol=0)
    assert any(s > c for s in sig) and any(s <= c for s in sig), "golden case must exercise both branches"
    for i, p in enumerate(proj):
        assert torch.equal(p, g[f"w{i}_proj"])
    torch.testing.assert_close(oa.mlp_forward(g["obs"], proj, b), g["mean_proj"], rtol=0, atol=1e-6)


def test_bf16_emulation_is_close_to_fp32(golden_dir):
    g = _load(golden_dir)
    w, b = _params(g)
    err = (oa.mlp_forward_bf16(g["obs"], w, b) - g["mean"]).abs().max().item()
    assert err < 2e-2, err                                     # tanh outputs in [-1, 1]; bf16 operands


def test_actor_noise_is_standard_normal():
    eps = oa.actor_noise(0x7AC0, np.arange(20000, dtype=np.int64), 3)
    assert abs(eps.mean().item()) < 0.02 and abs(eps.std().item() - 1.0) < 0.02
    assert math.isfinite(eps.abs().max().item())

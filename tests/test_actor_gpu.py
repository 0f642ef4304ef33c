"""Actor MLP kernels (through the C ABI / ActorMLP) against the oracle and the committed golden vectors
produced by the reference's own nets_asymmetry / ppo_asymmetry classes.  GPU only.

Bars:
  * FP32 CUDA-core path vs the float32 oracle (MLP.forward, nets_asymmetry.py:23-39): <= 2e-6 absolute on the tanh
    outputs (same operations; only the summation order of the dot products differs from the CPU BLAS);
  * tcgen05 bf16 path vs the bf16-operand / fp32-accumulate emulation of the same chain: <= 2e-3 absolute
    (accumulation order inside the tensor core), and <= 3e-2 vs the float32 oracle (bf16 operand rounding);
  * spectral norm: sigma within 1e-5 relative of torch.linalg.matrix_norm(ord=2), projected weights within 2e-6 relative;
  * sampling: action / log_p bit-level agreement with the oracle on identical Philox draws up to 1-ulp libm noise (2e-6).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "actor.npz"))
    g = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
    n = len([k for k in g if k.startswith("b")])
    return g, [g[f"w{i}"] for i in range(n)], [g[f"b{i}"] for i in range(n)]


def _random_mlp(sizes, seed, bias=0.1):
    gen = torch.Generator().manual_seed(seed)
    ws, bs = [], []
    for l in range(len(sizes) - 1):
        w = torch.empty(sizes[l + 1], sizes[l])
        torch.nn.init.orthogonal_(w, gain=(2 ** 0.5 if l + 2 < len(sizes) else 0.3), generator=gen)
        ws.append(w + 0.02 * torch.randn(w.shape, generator=gen))
        bs.append((torch.rand(sizes[l + 1], generator=gen) * 2 - 1) * bias)
    return ws, bs


def test_fp32_path_matches_reference_golden(golden_dir):
    from taco_b200 import ActorMLP
    g, w, b = _golden(golden_dir)
    a = ActorMLP(26, [w[0].shape[0], w[1].shape[0]], 4)
    a.load(w, b)
    mean = a.forward(g["obs"].cuda()).cpu()
    torch.testing.assert_close(mean, g["mean"], rtol=0, atol=2e-6)
    a.close()


def test_spectral_projection_matches_reference_golden(golden_dir):
    from taco_b200 import ActorMLP
    g, w, b = _golden(golden_dir)
    a = ActorMLP(26, [w[0].shape[0], w[1].shape[0]], 4)
    a.load(w, b, lipschitz_const=float(g["lipschitz"]))
    sig = torch.from_numpy(a.sigmas()).float()
    torch.testing.assert_close(sig, g["sigma_before"], rtol=1e-5, atol=0)
    for l in range(3):
        wl, bl = a.weights(l)
        ref = g[f"w{l}_proj"]
        assert np.abs(wl - ref.numpy()).max() <= 2e-6 * ref.abs().max().item() + 1e-9, l
        assert np.array_equal(bl, b[l].numpy())
    mean = a.forward(g["obs"].cuda()).cpu()
    torch.testing.assert_close(mean, g["mean_proj"], rtol=0, atol=3e-6)
    a.close()


@pytest.mark.parametrize("hidden,n", [([256, 256, 256], 4096 + 77), ([128, 64], 300), ([64], 128), ([256, 128, 192, 64], 1000)])
def test_fp32_and_tensor_core_paths_vs_oracle(hidden, n):
    from taco_b200 import ActorMLP
    from oracle import actor as oa
    sizes = [26] + hidden + [4]
    w, b = _random_mlp(sizes, seed=len(hidden) * 100 + n)
    obs = torch.randn(n, 1, 26, generator=torch.Generator().manual_seed(n)) * 0.7
    a = ActorMLP(26, hidden, 4)
    a.load(w, b)
    ref32 = oa.mlp_forward(obs, w, b)
    got32 = a.forward(obs.cuda()).cpu()
    assert (got32 - ref32).abs().max().item() <= 2e-6
    assert a.tensor_cores_available
    ref16 = oa.mlp_forward_bf16(obs, w, b)
    got16 = a.forward(obs.cuda(), tensor_cores=True).cpu()
    assert torch.isfinite(got16).all()
    err_emul = (got16 - ref16).abs().max().item()
    err_fp32 = (got16 - ref32).abs().max().item()
    assert err_emul <= 2e-3, (err_emul, err_fp32)
    assert err_fp32 <= 3e-2, (err_emul, err_fp32)
    # a second call on the same handle (persistent ring / barrier state is per launch) and a tile-unaligned tail
    got16b = a.forward(obs.cuda()[: n - 5], tensor_cores=True).cpu()
    assert torch.equal(got16b, got16[: n - 5])
    a.close()


def test_len_obs_5_input_width_130():
    """actor_input_dim = num_obs * len_obs (train_fpv_asymmetry_ppo.py:388-396): 3 K-chunks for the first layer."""
    from taco_b200 import ActorMLP
    from oracle import actor as oa
    hidden = [128, 128]
    w, b = _random_mlp([130] + hidden + [4], seed=5)
    obs = torch.randn(777, 5, 26, generator=torch.Generator().manual_seed(9)) * 0.5
    a = ActorMLP(130, hidden, 4)
    a.load(w, b)
    assert (a.forward(obs.cuda()).cpu() - oa.mlp_forward(obs, w, b)).abs().max().item() <= 2e-6
    got = a.forward(obs.cuda(), tensor_cores=True).cpu()
    assert (got - oa.mlp_forward_bf16(obs, w, b)).abs().max().item() <= 2e-3
    a.close()


@pytest.mark.parametrize("tc", [False, True])
def test_act_sampling_matches_oracle(tc):
    from taco_b200 import ActorMLP
    from oracle import actor as oa
    hidden, n, seed, step, off = [64, 64], 2048 + 3, 0x7AC0, 11, 5000
    w, b = _random_mlp([26] + hidden + [4], seed=3)
    log_std = torch.tensor([-0.5, -0.25, 0.0, 0.1])
    obs = torch.randn(n, 1, 26, generator=torch.Generator().manual_seed(1))
    a = ActorMLP(26, hidden, 4)
    a.load(w, b, log_std=log_std)
    action, clipped, logp, mean = [t.cpu() for t in a.act(obs.cuda(), step, seed=seed, env_offset=off, tensor_cores=tc)]
    eps = oa.actor_noise(seed, np.arange(off, off + n, dtype=np.int64), step)
    ref_action, ref_clipped, ref_logp = oa.act(mean, log_std, eps)       # the kernel's own mean: isolates the sampling
    torch.testing.assert_close(action, ref_action, rtol=0, atol=5e-6)
    torch.testing.assert_close(logp, ref_logp, rtol=0, atol=2e-5)
    assert torch.equal(clipped, action.clamp(-1, 1))
    ref_mean = oa.mlp_forward_bf16(obs, w, b) if tc else oa.mlp_forward(obs, w, b)
    assert (mean - ref_mean).abs().max().item() <= (2e-3 if tc else 2e-6)
    a.close()


def test_shape_errors_are_loud():
    from taco_b200 import ActorMLP
    a = ActorMLP(26, [100], 4)                    # 100 is not a multiple of 64: FP32 path only
    w, b = _random_mlp([26, 100, 4], seed=1)
    a.load(w, b)
    assert not a.tensor_cores_available
    obs = torch.zeros(8, 26, device="cuda")
    a.forward(obs)
    with pytest.raises(RuntimeError, match="tensor-core path unavailable"):
        a.forward(obs, tensor_cores=True)
    with pytest.raises(ValueError):
        a.forward(torch.zeros(8, 27, device="cuda"))
    a.close()


def test_device_side_projection_of_torch_parameters(golden_dir):
    """spectral_normalize_ == PPO.spectral_normalize_actors on nn.Linear weights that live on the GPU (golden: the reference's own
    method), in place, biases untouched."""
    from taco_b200 import spectral_normalize_
    g, w, b = _golden(golden_dir)
    lin = [torch.nn.Linear(x.shape[1], x.shape[0]).cuda() for x in w]
    with torch.no_grad():
        for m, wl, bl in zip(lin, w, b):
            m.weight.copy_(wl); m.bias.copy_(bl)
    params = [p for m in lin for p in m.parameters()]
    sig = spectral_normalize_(params, float(g["lipschitz"]))
    torch.testing.assert_close(sig.cpu().float(), g["sigma_before"], rtol=1e-5, atol=0)
    for l, m in enumerate(lin):
        ref = g[f"w{l}_proj"]
        assert (m.weight.detach().cpu() - ref).abs().max().item() <= 2e-6 * ref.abs().max().item() + 1e-9, l
        assert torch.equal(m.bias.detach().cpu(), b[l])

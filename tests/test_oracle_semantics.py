"""Freeze the torch-CPU semantics the CUDA kernel is written against (SURVEY.md section 8c "verified semantics"
plus the ones found while chasing bit parity).  CPU only."""
import numpy as np
import torch


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def test_negative_index_wraps_and_overlapping_shift_is_memmove():
    buf = torch.arange(100.0).repeat(2, 4, 1)
    idx = torch.clamp(torch.tensor([0, 5]) - 1, max=3)               # fpv_asymmetry.py:366 with len = 0 -> index -1
    got = buf[torch.arange(2), :, idx]
    assert got[0, 0].item() == 99.0 and got[1, 0].item() == 3.0
    buf[:, :, 0:-10] = buf[:, :, 10:]                                  # fpv_asymmetry.py:378
    assert torch.equal(buf[0, 0, :90], torch.arange(10.0, 100.0)) and torch.equal(buf[0, 0, 90:], torch.arange(90.0, 100.0))


def test_sum_is_sequential_and_norm_accumulates_with_fma():
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(50000, 4, generator=g) * 1000).float()
    a = x.numpy()
    assert np.array_equal(torch.sum(x, dim=1).numpy(), ((a[:, 0] + a[:, 1]) + a[:, 2]) + a[:, 3])
    v = ((torch.rand(50000, 3, generator=g) * 10) - 5).float()
    b = v.numpy().astype(np.float64)
    acc = _f32(b[:, 0] * b[:, 0]).astype(np.float64)
    acc = _f32(acc + b[:, 1] * b[:, 1]).astype(np.float64)            # fma(y, y, acc): product exact in float64
    n2 = _f32(np.sqrt(_f32(acc).astype(np.float64)))
    assert np.array_equal(torch.norm(v[:, :2], dim=1).numpy(), n2)
    acc = _f32(acc + b[:, 2] * b[:, 2]).astype(np.float64)
    assert np.array_equal(torch.norm(v, dim=1).numpy(), _f32(np.sqrt(acc)))


def test_python_scalar_over_tensor_is_reciprocal_times_scalar():
    t = torch.tensor([0.017, 0.0163, 0.0177, 0.001], dtype=torch.float32)
    got = (0.001 / t).numpy()                                          # thrust_dynamics.py:84
    want = (np.float32(1.0) / t.numpy()) * np.float32(0.001)
    assert np.array_equal(got, want)
    assert not np.array_equal(got, np.float32(0.001) / t.numpy())      # differs from a true division by 1 ulp somewhere


def test_pow_3_is_repeated_multiplication_and_mvn_std():
    x = torch.tensor([1.2345678, 0.3333333, 7.7777], dtype=torch.float32)
    assert np.array_equal((x ** 3).numpy(), (x * x * x).numpy()) and np.array_equal((x ** 2).numpy(), (x * x).numpy())
    # nets_asymmetry.py:338-339: scale_tril = diag(exp(log_std)^2)  =>  std = exp(2 log_std)
    log_std = torch.tensor([0.3, -0.2])
    cov = torch.diag(log_std.exp() * log_std.exp())
    dist = torch.distributions.MultivariateNormal(torch.zeros(2), scale_tril=cov)
    assert torch.allclose(dist.stddev, torch.exp(2 * log_std))

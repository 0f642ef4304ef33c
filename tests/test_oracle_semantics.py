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


def test_fma32_is_an_exact_float32_fma():
    """oracle.rigid_body.fma32 against exact rational arithmetic, on inputs built to cancel (where double rounding bites)."""
    from fractions import Fraction
    import numpy as np
    from oracle.rigid_body import fma32
    rng = np.random.default_rng(0)
    a = rng.standard_normal(1500).astype(np.float32)
    b = rng.standard_normal(1500).astype(np.float32)
    c = (-(a.astype(np.float64) * b.astype(np.float64))).astype(np.float32) + rng.standard_normal(1500).astype(np.float32) * np.float32(1e-7)
    c[::3] = rng.standard_normal(500).astype(np.float32)
    r = fma32(a, b, c)
    for a_, b_, c_, r_ in zip(a, b, c, r):
        ex = Fraction(float(a_)) * Fraction(float(b_)) + Fraction(float(c_))
        f = np.float32(float(ex))
        best = min([f, np.nextafter(f, np.float32(np.inf)), np.nextafter(f, np.float32(-np.inf))], key=lambda x: abs(Fraction(float(x)) - ex))
        assert best == r_


def test_c_integrator_equals_its_numpy_twin():
    """oracle/rigid_body.c (the executable specification of the integrator, C fmaf) == _integrate_py (exact FMA emulation)."""
    import torch
    from oracle import rigid_body as rb
    assert rb._c_lib() is not None, "oracle/librigid_body.so missing: run __graft_entry__.build() (needs gcc)"
    g = torch.Generator().manual_seed(1)
    n = 6000
    pos = torch.randn(n, 3, generator=g) * 3
    q = torch.randn(n, 4, generator=g); q = q / q.norm(dim=1, keepdim=True)
    v, w = torch.randn(n, 3, generator=g) * 4, torch.randn(n, 3, generator=g) * 20
    fb, tb = torch.randn(n, 3, generator=g) * 5, torch.randn(n, 3, generator=g) * 0.05
    fb[:100] = 0; tb[:100] = 0                                  # the zero-wrench step after a reset
    for sub in (1, 2, 3):
        a = rb.integrate(pos, q, v, w, fb, tb, 0.001, sub, use_c=True)
        b = rb.integrate(pos, q, v, w, fb, tb, 0.001, sub, use_c=False)
        assert all(torch.equal(x, y) for x, y in zip(a, b))
        assert torch.all((a[1].norm(dim=1) - 1).abs() < 1e-6)

"""Critic kernels (LSTM encoder + MLP -> value, through the C ABI / CriticLSTM) against the oracle and the committed golden
vectors produced by the reference's own nets_asymmetry classes.  GPU only.

Bars:
  * FP32 CUDA-core path vs the golden values of the reference (LSTMEncoder + MLP, nets_asymmetry.py:128-136,:23-39,:350-352):
    <= 5e-6 absolute (same operations; the summation order of the dot products differs from the CPU BLAS and expf / tanhf
    are CUDA's, 1-2 ulp from torch's);
  * tcgen05 bf16 path vs the bf16-operand / fp32-accumulate emulation of the same chain (oracle/critic.py:
    critic_forward_bf16): <= 1.5e-2 absolute on O(1) values (measured <= 8e-3) -- the tensor core's accumulation order plus MUFU.TANH for the five
    gate non-linearities (2^-11 relative each, applied 5 steps deep) -- and <= 6e-2 vs the float32 oracle (bf16 operands).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(golden_dir, tag):
    from test_critic_oracle import load_case
    return load_case(golden_dir, tag)


def _random_critic(in_dim, hid, mlp_hidden, seed, nl=1):
    gen = torch.Generator().manual_seed(seed)
    lstm = []
    for l in range(nl):
        in_l = in_dim if l == 0 else hid
        w_ih = torch.empty(4 * hid, in_l); w_hh = torch.empty(4 * hid, hid)
        torch.nn.init.xavier_uniform_(w_ih, generator=gen); torch.nn.init.xavier_uniform_(w_hh, generator=gen)   # LSTMEncoder.para_init
        lstm.append((w_ih * 1.5, w_hh * 1.5, (torch.rand(4 * hid, generator=gen) - 0.5) * 0.6, (torch.rand(4 * hid, generator=gen) - 0.5) * 0.6))
    sizes = [hid] + list(mlp_hidden) + [1]
    ws, bs = [], []
    for l in range(len(sizes) - 1):
        w = torch.empty(sizes[l + 1], sizes[l])
        torch.nn.init.orthogonal_(w, gain=(2 ** 0.5 if l + 2 < len(sizes) else 6.0), generator=gen)   # values of O(1)
        ws.append(w); bs.append((torch.rand(sizes[l + 1], generator=gen) - 0.5) * 0.2)
    return lstm, ws, bs


@pytest.mark.parametrize("tag", ["a", "b"])          # a: 1 LSTM layer x 48, MLP 64-32;  b: 2 LSTM layers x 32, MLP 40
def test_fp32_path_matches_reference_golden(golden_dir, tag):
    from taco_b200 import CriticLSTM
    g, lstm, w, b = _case(golden_dir, tag)
    hid = lstm[0][1].shape[1]
    c = CriticLSTM(26, 5, hid, [x.shape[0] for x in w[:-1]], lstm_layers=len(lstm))
    c.load(lstm, w, b)
    value = c.forward(g["states"].cuda()).cpu()
    torch.testing.assert_close(value, g["value"], rtol=0, atol=5e-6)
    c.close()


@pytest.mark.parametrize("n", [1, 31, 33, 1000])
def test_fp32_path_ragged_sizes_vs_oracle(n):
    from taco_b200 import CriticLSTM
    from oracle import critic as oc
    lstm, w, b = _random_critic(26, 40, [72, 24], seed=n)
    states = torch.randn(n, 5, 26, generator=torch.Generator().manual_seed(100 + n)) * 1.2
    c = CriticLSTM(26, 5, 40, [72, 24])
    c.load(lstm, w, b)
    torch.testing.assert_close(c.forward(states.cuda()).cpu(), oc.critic_forward(states, lstm, w, b), rtol=0, atol=5e-6)
    c.close()


@pytest.mark.parametrize("hid,mlp_hidden", [(64, [256, 256, 256]), (64, [128]), (32, [64, 192]), (48, [256, 64]), (16, [64])])
@pytest.mark.parametrize("n", [128, 300, 40_000])
def test_tensor_core_path(hid, mlp_hidden, n):
    from taco_b200 import CriticLSTM
    from oracle import critic as oc
    lstm, w, b = _random_critic(26, hid, mlp_hidden, seed=hid + len(mlp_hidden))
    states = torch.randn(n, 5, 26, generator=torch.Generator().manual_seed(7 + n)) * 1.2
    c = CriticLSTM(26, 5, hid, mlp_hidden)
    assert c.tensor_cores_available
    c.load(lstm, w, b)
    v_tc = c.forward(states.cuda(), tensor_cores=True).cpu()
    v_fp = c.forward(states.cuda(), tensor_cores=False).cpu()
    m = min(n, 4096)                                      # the CPU emulation is slow: check a prefix and a suffix of the rows
    for sl in (slice(0, m), slice(n - m, n)):
        emu = oc.critic_forward_bf16(states[sl], lstm, w, b)
        assert float(emu.abs().mean()) > 0.05
        err = (v_tc[sl] - emu).abs().max().item()
        print(f"critic tc vs bf16 emulation: hid={hid} mlp={mlp_hidden} n={n} max|err|={err:.2e} mean|v|={float(emu.abs().mean()):.3f}")
        assert err <= 1.5e-2
    err32 = (v_tc - v_fp).abs().max().item()
    print(f"critic tc vs fp32 kernel: max|err|={err32:.2e}")
    assert err32 <= 6e-2
    assert torch.isfinite(v_tc).all()
    c.close()


@pytest.mark.parametrize("in_dim,seq_len,hid", [(26, 1, 64), (26, 8, 32), (30, 5, 64), (2, 3, 16), (26, 5, 64)])
def test_tensor_core_path_other_sequence_shapes(in_dim, seq_len, hid):
    """Edge shapes of the tcgen05 path: a single frame (no x restaging), the longest supported history, the widest / narrowest
    even input (the two bias columns sit right after the features), against the FP32 kernel and the bf16 emulation."""
    from taco_b200 import CriticLSTM
    from oracle import critic as oc
    lstm, w, b = _random_critic(in_dim, hid, [64], seed=in_dim * 100 + seq_len)
    n = 777
    states = torch.randn(n, seq_len, in_dim, generator=torch.Generator().manual_seed(in_dim + seq_len)) * 1.2
    c = CriticLSTM(in_dim, seq_len, hid, [64])
    assert c.tensor_cores_available
    c.load(lstm, w, b)
    v_tc = c.forward(states.cuda(), tensor_cores=True).cpu()
    assert (v_tc - oc.critic_forward_bf16(states, lstm, w, b)).abs().max().item() <= 1.5e-2
    torch.testing.assert_close(c.forward(states.cuda()).cpu(), oc.critic_forward(states, lstm, w, b), rtol=0, atol=5e-6)
    c.close()


def test_fp32_path_deep_lstm_vs_oracle():
    """Three stacked LSTM layers, odd widths, a sequence of 7 frames of 11 features: only the FP32 kernel takes this shape."""
    from taco_b200 import CriticLSTM
    from oracle import critic as oc
    lstm, w, b = _random_critic(11, 37, [19, 5], seed=77, nl=3)
    states = torch.randn(203, 7, 11, generator=torch.Generator().manual_seed(5))
    c = CriticLSTM(11, 7, 37, [19, 5], lstm_layers=3)
    assert not c.tensor_cores_available
    c.load(lstm, w, b)
    torch.testing.assert_close(c.forward(states.cuda()).cpu(), oc.critic_forward(states, lstm, w, b), rtol=0, atol=5e-6)
    c.close()


def test_tensor_core_path_rejects_unsupported_shapes():
    from taco_b200 import CriticLSTM
    c = CriticLSTM(26, 5, 128, [256])                      # LSTM width 128: the cell state would not fit the register file
    assert not c.tensor_cores_available
    lstm, w, b = _random_critic(26, 128, [256], seed=3)
    c.load(lstm, w, b)
    with pytest.raises(RuntimeError, match="tensor-core path unavailable"):
        c.forward(torch.zeros(8, 5, 26, device="cuda"), tensor_cores=True)
    assert torch.isfinite(c.forward(torch.zeros(8, 5, 26, device="cuda"))).all()
    with pytest.raises(ValueError):
        c.forward(torch.zeros(8, 4, 26, device="cuda"))
    c.close()
    with pytest.raises(RuntimeError, match="mlp_sizes"):
        from taco_b200 import _capi
        import ctypes as C
        arr = (C.c_int32 * 2)(32, 1)
        h = C.c_void_p()
        _capi.check(_capi.lib().taco_critic_create(0, 26, 5, 64, 1, arr, 2, C.byref(h)), "taco_critic_create")

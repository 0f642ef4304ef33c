"""Pin oracle/critic.py to the reference's own nets_asymmetry classes (LSTMEncoder, MLP, the critic branch of
PPO_ActorCritic.act) through tests/golden/critic.npz (oracle/make_golden.py: critic()).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import critic as oc


def load_case(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, "critic.npz"))
    g = {k[len(tag) + 1:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith(tag + "_")}
    nl = len([k for k in g if k.startswith("weight_ih_l")])
    lstm = [(g[f"weight_ih_l{l}"], g[f"weight_hh_l{l}"], g[f"bias_ih_l{l}"], g[f"bias_hh_l{l}"]) for l in range(nl)]
    nm = len([k for k in g if k.startswith("b") and k[1:].isdigit()])
    return g, lstm, [g[f"w{i}"] for i in range(nm)], [g[f"b{i}"] for i in range(nm)]


@pytest.mark.parametrize("tag", ["a", "b"])          # a: 1 LSTM layer x 48, MLP 64-32;  b: 2 LSTM layers x 32, MLP 40
def test_critic_vs_reference(golden_dir, tag):
    g, lstm, w, b = load_case(golden_dir, tag)
    enc = oc.lstm_last_hidden(g["states"], lstm)
    torch.testing.assert_close(enc, g["enc"], rtol=0, atol=1e-6)          # same arithmetic; torch's fused LSTM cell orders the four adds differently
    value = oc.critic_forward(g["states"], lstm, w, b)
    torch.testing.assert_close(value, g["value"], rtol=0, atol=3e-6)
    assert torch.equal(g["act_value"], g["value"])                          # act() runs exactly encoder -> mlp (nets_asymmetry.py:350-352)
    assert value.shape == (g["states"].shape[0], 1) and float(value.abs().mean()) > 0.05   # the case is not degenerate


@pytest.mark.parametrize("tag", ["a", "b"])
def test_bf16_emulation_is_close_to_fp32(golden_dir, tag):
    g, lstm, w, b = load_case(golden_dir, tag)
    err = (oc.critic_forward_bf16(g["states"], lstm, w, b) - g["value"]).abs().max().item()
    assert err < 5e-2, err

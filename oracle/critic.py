"""Oracle for the critic in the rollout loop.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates, in plain torch float32 on the CPU (explicit gate arithmetic, no nn.LSTM):
  * LSTMEncoder.forward               IsaacGymEnvs/algorithms/nets_asymmetry.py:128-136
    (nn.LSTM(input, hidden, num_layers, batch_first=True) over the (N, len_states, 26) state history, zero initial state;
     the encoder output is the top layer's hidden state after the LAST time step, `x[:, -1, :]`)
  * MLP.forward with Identity output  IsaacGymEnvs/algorithms/nets_asymmetry.py:23-39, :318
  * PPO_ActorCritic.act, critic branch IsaacGymEnvs/algorithms/nets_asymmetry.py:350-352
Gate order of torch's LSTM weights: rows [0,H) input gate i, [H,2H) forget gate f, [2H,3H) cell candidate g, [3H,4H) output
gate o;  c' = sigmoid(f) * c + sigmoid(i) * tanh(g),  h' = sigmoid(o) * tanh(c').
Pinned against the reference's own nets_asymmetry classes by tests/golden/critic.npz (oracle/make_golden.py: critic()).
"""
import torch


def lstm_last_hidden(x, layers):
    """x (N, T, In) float32; layers = [(w_ih (4H,In), w_hh (4H,H), b_ih (4H), b_hh (4H)), ...] bottom layer first.
    Returns the top layer's h after the last time step, (N, H)."""
    n, t_len, _ = x.shape
    seq = x
    for (w_ih, w_hh, b_ih, b_hh) in layers:
        hid = w_hh.shape[1]
        h = torch.zeros(n, hid, dtype=x.dtype)
        c = torch.zeros(n, hid, dtype=x.dtype)
        outs = []
        for t in range(t_len):
            gates = torch.nn.functional.linear(seq[:, t, :], w_ih, b_ih) + torch.nn.functional.linear(h, w_hh, b_hh)
            i, f, g, o = gates.chunk(4, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        seq = torch.stack(outs, dim=1)
    return seq[:, -1, :]


def mlp_identity(x, weights, biases):
    """nets_asymmetry.py:37-39 with activation=ReLU, output_activation=Identity (:318)."""
    x = x.contiguous().view(x.size(0), -1)
    n = len(weights)
    for l in range(n):
        x = torch.nn.functional.linear(x, weights[l], biases[l])
        if l + 1 < n:
            x = torch.relu(x)
    return x


def critic_forward(states, lstm_layers, weights, biases):
    """value (N, 1) = critic_mlp(critic_encoder(states)), nets_asymmetry.py:350-352."""
    return mlp_identity(lstm_last_hidden(states, lstm_layers), weights, biases)


def _bf(x):
    return x.to(torch.bfloat16).double()


def critic_forward_bf16(states, lstm_layers, weights, biases, cell_dtype=torch.float32):
    """What the tensor-core kernel computes: every matrix product takes bf16 operands (inputs, hidden state, weights) with
    fp32 accumulation; biases, gate non-linearities, the cell state and ReLU are fp32; h and the MLP activations are rounded
    to bf16 only as the next product's operand."""
    n, t_len, _ = states.shape
    seq = states.float()
    for (w_ih, w_hh, b_ih, b_hh) in lstm_layers:
        hid = w_hh.shape[1]
        h = torch.zeros(n, hid)
        c = torch.zeros(n, hid)
        outs = []
        wi, wh = _bf(w_ih), _bf(w_hh)
        bias = (b_ih + b_hh).float()
        for t in range(t_len):
            gates = (_bf(seq[:, t, :]) @ wi.t() + _bf(h) @ wh.t()).float() + bias
            i, f, g, o = gates.chunk(4, dim=1)
            c = (torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)).to(cell_dtype).float()
            h = torch.sigmoid(o) * torch.tanh(c)
            outs.append(h)
        seq = torch.stack(outs, dim=1)
    x = seq[:, -1, :]
    nl = len(weights)
    for l in range(nl):
        x = (_bf(x) @ _bf(weights[l]).t()).float() + biases[l]
        if l + 1 < nl:
            x = torch.relu(x)
    return x

"""Oracle for the actor MLP in the rollout loop.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates, in plain torch float32 on the CPU:
  * MLP.forward                       IsaacGymEnvs/algorithms/nets_asymmetry.py:23-39
    ([Linear -> ReLU] x L -> Linear -> Tanh on the flattened (N, len_obs*26) observation)
  * PPO_ActorCritic.act, actor branch IsaacGymEnvs/algorithms/nets_asymmetry.py:326-346
    (scale_tril = diag(exp(log_std)^2), so the standard deviation is exp(2 log_std); sample; log_prob)
  * PPO.spectral_normalize_actors     IsaacGymEnvs/algorithms/ppo_asymmetry.py:398-404
  * the clip PPO applies before env.step, ppo_asymmetry.py:310
Pinned against the reference's own nets_asymmetry module by tests/golden/actor.npz
(oracle/make_golden.py: actor()).  The noise is injected (eps) so that the CUDA kernels and the
oracle consume identical Philox draws; the reference samples from torch's global generator.
"""
import math

import numpy as np
import torch

from . import philox as px

STREAM_ACTOR = 6


def mlp_forward(obs, weights, biases):
    """nets_asymmetry.py:37-39 with activation=ReLU, output_activation=Tanh (:313)."""
    x = obs.contiguous().view(obs.size(0), -1)
    n = len(weights)
    for l in range(n):
        x = torch.nn.functional.linear(x, weights[l], biases[l])
        x = torch.relu(x) if l + 1 < n else torch.tanh(x)
    return x


def mlp_forward_bf16(obs, weights, biases):
    """What the tcgen05 kernel computes: bf16 operands (activations and weights of every layer, the output layer
    included), fp32 accumulation, fp32 bias, ReLU in fp32, then rounding to bf16 for the next layer's operand."""
    x = obs.contiguous().view(obs.size(0), -1).float()
    n = len(weights)
    for l in range(n):
        xa = x.to(torch.bfloat16).double()
        wa = weights[l].to(torch.bfloat16).double()
        x = (xa @ wa.t()).float() + biases[l]
        x = torch.relu(x) if l + 1 < n else torch.tanh(x)
    return x


def spectral_normalize(weights, lipschitz_const):
    """ppo_asymmetry.py:398-404: hard projection of every >= 2-D parameter; returns (weights, sigmas)."""
    out, sig = [], []
    for w in weights:
        w = w.clone()
        s = torch.linalg.matrix_norm(w, ord=2)
        sig.append(float(s))
        if lipschitz_const > 0 and s > lipschitz_const:
            w *= lipschitz_const / s
        out.append(w)
    return out, sig


def actor_noise(seed, gid, step):
    """eps (N,4) ~ N(0,1): Box-Muller over the Philox words of (gid, step, 0, STREAM_ACTOR), same pairing as
    taco_b200/csrc/actor_tc.cuh: actor_tail."""
    r = px.draw(seed, gid, step, 0, STREAM_ACTOR)
    z0, z1 = px.box_muller(r[:, 0], r[:, 1])
    z2, z3 = px.box_muller(r[:, 2], r[:, 3])
    return torch.from_numpy(np.stack([z0, z1, z2, z3], axis=1).astype(np.float32))


def act(mean, log_std, eps):
    """nets_asymmetry.py:336-346 with injected noise + ppo_asymmetry.py:310.  Returns action, clipped, logp."""
    e = log_std.exp()
    std = e * e                                             # scale_tril = diag(exp(log_std) * exp(log_std))
    k = mean.shape[1]
    action = mean + std[:k] * eps[:, :k]
    clipped = action.clamp(-1.0, 1.0)
    logp = -0.5 * (eps[:, :k] ** 2).sum(dim=1) - std[:k].log().sum() - 0.5 * k * math.log(2.0 * math.pi)
    return action, clipped, logp

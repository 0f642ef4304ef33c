"""Oracle for the rollout buffer's return / advantage pass.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates, in plain torch float32 on the CPU:
  * PPOReplayBuffer.compute_returns_and_advantage   IsaacGymEnvs/algorithms/buffer_asymmetry.py:93-132
    (backward GAE(lambda) recursion over the horizon, ret = adv + value, then
     adv = (adv - adv.mean()) / (adv.std() + 1e-8) over ALL (horizon x envs) samples, unbiased std)
  * the time-out bootstrap PPO applies to the reward before storing it
    IsaacGymEnvs/algorithms/ppo_asymmetry.py:313-324: rewards[truncated] += gamma * V(obs, states) where obs / states are
    the PRE-step observations, i.e. the very `value` stored for that step
Pinned against the reference's own PPOReplayBuffer class by tests/golden/gae.npz
(oracle/make_golden.py: gae()).
"""
import torch


def bootstrap_timeouts(rew, value, done, time_outs, gamma):
    """ppo_asymmetry.py:313-324 without the host round trip: truncated = time_outs * dones (non-zero)."""
    trunc = (time_outs.float() * done.float()) != 0
    return torch.where(trunc, rew + gamma * value, rew)


def gae(rew, done, value, last_value, gamma, lam):
    """buffer_asymmetry.py:112-127.  rew / done / value: (H, N, 1) float32; last_value: (N, 1).
    Returns (adv, ret) BEFORE normalisation."""
    H = rew.shape[0]
    adv = torch.zeros_like(rew)
    last_gae_lam = 0
    for step in reversed(range(H)):
        next_values = last_value if step == H - 1 else value[step + 1]
        next_non_terminal = 1.0 - done[step].float()
        td_target = rew[step] + next_non_terminal * gamma * next_values
        delta = td_target - value[step]
        last_gae_lam = delta + next_non_terminal * gamma * lam * last_gae_lam
        adv[step] = last_gae_lam
    return adv, adv + value


def normalize(adv):
    """buffer_asymmetry.py:132."""
    return (adv - adv.mean()) / (adv.std() + 1e-8)


def moments(adv):
    """[sum, sum of squares, count] in float64: the vector a multi-GPU job all-reduces."""
    a = adv.double().flatten()
    return torch.stack([a.sum(), (a * a).sum(), torch.tensor(float(a.numel()), dtype=torch.float64)])


def normalize_from_moments(adv, m):
    """What the CUDA normalisation computes from the (all-reduced) moments: mean and unbiased std in float64, rounded
    to float32, then the reference expression in float32."""
    s, ss, n = float(m[0]), float(m[1]), float(m[2])
    mean = s / n
    var = max(ss - n * mean * mean, 0.0) / max(n - 1.0, 1.0)
    mean32 = torch.tensor(mean, dtype=torch.float32)
    std32 = torch.tensor(var ** 0.5, dtype=torch.float32)
    return (adv - mean32) / (std32 + 1e-8)

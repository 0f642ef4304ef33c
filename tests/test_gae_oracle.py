"""Pin oracle/gae.py to the reference's own PPOReplayBuffer (buffer_asymmetry.py:93-132) and to the time-out bootstrap of
ppo_asymmetry.py:313-324 through tests/golden/gae.npz (oracle/make_golden.py: gae()).  CPU only."""
import os

import numpy as np
import torch

from oracle import gae as og


def _load(golden_dir):
    z = np.load(os.path.join(golden_dir, "gae.npz"))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def test_bootstrap_vs_reference(golden_dir):
    g = _load(golden_dir)
    aug = og.bootstrap_timeouts(g["rew"], g["value"].squeeze(-1), g["done"], g["time_outs"], float(g["gamma"]))
    assert torch.equal(aug, g["rew_aug"])
    assert (g["time_outs"].bool() & (g["done"] != 0)).any(), "golden case must contain truncated steps"


def test_gae_and_normalisation_vs_reference(golden_dir):
    g = _load(golden_dir)
    adv, ret = og.gae(g["rew_aug"].unsqueeze(-1), g["done"].unsqueeze(-1), g["value"], g["last_value"], float(g["gamma"]), float(g["lam"]))
    assert torch.equal(ret, g["ret"])                       # bit-exact: same float32 operations in the same order
    assert torch.equal(og.normalize(adv), g["adv_norm"])


def test_moment_normalisation_matches_torch_mean_std(golden_dir):
    """The multi-GPU form (all-reduced [sum, sumsq, n] in float64) agrees with adv.mean() / adv.std() to float32 rounding."""
    g = _load(golden_dir)
    adv, _ = og.gae(g["rew_aug"].unsqueeze(-1), g["done"].unsqueeze(-1), g["value"], g["last_value"], float(g["gamma"]), float(g["lam"]))
    got = og.normalize_from_moments(adv, og.moments(adv))
    torch.testing.assert_close(got, g["adv_norm"], rtol=0, atol=2e-6)
    # moments add across shards
    m = og.moments(adv[:, :30]) + og.moments(adv[:, 30:])
    torch.testing.assert_close(m, og.moments(adv), rtol=1e-13, atol=0)

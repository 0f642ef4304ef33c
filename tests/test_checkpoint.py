"""Checkpoint / export paths around the native rollout (SURVEY.md section 5 "checkpoint / resume"; ppo_asymmetry.py:369-393,
452-468): agent state dicts interchangeable with the reference's parameter names, TorchScript actor export, and the env-state
checkpoint (which the reference does not have) continuing bit-identically."""
import os

import numpy as np
import pytest
import torch


def _agent():
    from taco_b200.ppo import TorchActorCritic
    torch.manual_seed(3)
    return TorchActorCritic(26, 4, [64, 64], 26, 32, [64], lstm_layers=1)


def test_state_dict_round_trip_and_reference_parameter_names(tmp_path):
    from taco_b200.ppo import load_checkpoint, make_optimizer, save_checkpoint, PPOConfig
    a = _agent()
    names = set(a.state_dict())
    # the names nets_asymmetry.PPO_ActorCritic produces for this configuration (MLP.layers Sequential, nn.LSTM under .layers)
    assert {"log_std", "actor_mlp.layers.0.weight", "actor_mlp.layers.4.bias", "critic_encoder.layers.weight_ih_l0",
            "critic_encoder.layers.bias_hh_l0", "critic_mlp.layers.2.weight"} <= names
    p = str(tmp_path / "model_5_0.1.pt")
    save_checkpoint(p, a, para_only=True)                       # PPO.save(para_only=True): a bare state dict
    b = _agent()
    with torch.no_grad():
        for prm in b.parameters():
            prm.add_(1.0)
    assert load_checkpoint(p, b) is None
    for (k, v), (_, w) in zip(a.state_dict().items(), b.state_dict().items()):
        assert torch.equal(v, w), k
    # resumable form: optimiser moments and epoch come back too
    opt = make_optimizer(a, PPOConfig())
    a.evaluate(torch.randn(8, 1, 26), torch.randn(8, 5, 26), torch.randn(8, 4))[0].sum().backward()
    opt.step()
    p2 = str(tmp_path / "resume.pt")
    save_checkpoint(p2, a, optimizer=opt, epoch=17, para_only=False)
    c = _agent()
    opt_c = make_optimizer(c, PPOConfig())
    assert load_checkpoint(p2, c, optimizer=opt_c) == 17
    sa, sc = opt.state_dict()["state"], opt_c.state_dict()["state"]
    assert sa.keys() == sc.keys() and all(torch.equal(sa[k]["exp_avg_sq"], sc[k]["exp_avg_sq"]) for k in sa)


def test_reference_state_dict_loads_when_the_reference_is_present(tmp_path):
    """Build container only: a state dict saved by the reference's own PPO_ActorCritic loads into TorchActorCritic and gives the
    same action mean and value."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present")
    ns = ref_loader.load()
    torch.manual_seed(0)
    para = {"actor_critic_mlp_dict": {"actor_input_dim": 26, "actor_output_dim": 4, "critic_input_dim": 26, "critic_output_dim": 1,
                                      "actor_hidden_sizes": [64, 64], "critic_hidden_sizes": [64], "activation": torch.nn.ReLU},
            "use_actor_encoder": False, "use_critic_encoder": True, "share_encoder": False, "critic_encoder_type": "LSTM",
            "critic_encoder_dict": {"encoder_type": "LSTM", "input_size": 26, "output_size": 32, "num_layers": 1, "bidirectional": False}}
    try:
        ref = ns.nets.PPO_ActorCritic(para)
    except Exception as exc:                                   # constructor dict keys differ between reference revisions
        pytest.skip(f"reference PPO_ActorCritic not constructible with the reconstructed dict: {exc!r}")
    p = str(tmp_path / "ref_model.pt")
    torch.save(ref.state_dict(), p)
    from taco_b200.ppo import load_checkpoint
    ours = _agent()
    load_checkpoint(p, ours)
    obs, st = torch.randn(16, 1, 26), torch.randn(16, 5, 26)
    with torch.no_grad():
        _, _, v_ref, m_ref, _ = ref.act(obs, st, deterministic=True)
        m = ours(obs)
        v = ours.critic_mlp(ours.critic_encoder(st))
    assert torch.allclose(m, m_ref, atol=1e-6) and torch.allclose(v, v_ref, atol=1e-6)


def test_export_actor_torchscript(tmp_path):
    from taco_b200.ppo import export_actor
    a = _agent()
    p = str(tmp_path / "actor_0.pt")
    eager, traced = export_actor(a, p, 1, 26, device="cpu")
    assert torch.equal(eager, traced)
    mod = torch.jit.load(p)
    x = torch.randn(5, 1, 26)
    with torch.no_grad():
        assert torch.allclose(mod(x), a(x), atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("task,dr", [("flip", False), ("mix", True)])
def test_env_state_checkpoint_continues_bit_identically(task, dr, tmp_path):
    import taco_b200
    cfg = taco_b200.make_cfg(task, 1000, domain_randomization=dr, observation_noise=dr, **{"env.maxEpisodeLength": 12})
    a = taco_b200.FpvVecTask(cfg, seed=11)
    a.reset()
    for t in range(9):
        a.step(a.random_actions(t))
    a.difficulty = 0.6
    path = str(tmp_path / "env.npy")
    a.save_state(path)
    want = []
    for t in range(9, 16):
        o, r, x, e = a.step(a.random_actions(t))
        want.append((o["obs"].clone(), o["states"].clone(), r.clone(), x.clone(), e["time_outs"].clone()))
    stats_a = a.stats().cpu()
    b = taco_b200.FpvVecTask(cfg, seed=999)                     # other seed / difficulty: both come from the checkpoint
    b.load_state(path)
    assert b.step_count == 9 and b.difficulty == pytest.approx(0.6)
    for t, w in zip(range(9, 16), want):
        o, r, x, e = b.step(b.random_actions(t))
        got = (o["obs"], o["states"], r, x, e["time_outs"])
        for g_, w_ in zip(got, w):
            assert torch.equal(g_, w_), t
    assert torch.equal(b.stats().cpu(), stats_a)
    assert np.array_equal(a.export_state(), b.export_state())
    with pytest.raises(RuntimeError):                           # a checkpoint of another configuration is refused
        c = taco_b200.FpvVecTask(taco_b200.make_cfg(task, 1001), seed=1)
        c.load_state_checkpoint(np.load(path))
    a.close(); b.close()

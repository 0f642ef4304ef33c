"""taco_b200.ppo (TorchActorCritic, schedules, ppo_update) against the reference's own ``PPO.update`` / ``PPO_ActorCritic``
(tests/golden/ppo_update.npz, produced by oracle/make_golden.py: ppo_update() from ppo_asymmetry.py:137-258 run on a fixed
buffer with fixed minibatch indices).  Bars: parameters after the update (6 Adam steps, gradient clipping, schedules) within
2e-5 absolute of the reference's, logged scalars within 1e-5 relative.  The CPU tests run everywhere; the GPU test uses the
device-side spectral projection (taco_spectral_project) instead of the reference's SVD."""
import os

import numpy as np
import pytest
import torch


def _load(golden_dir, tag):
    z = np.load(os.path.join(golden_dir, "ppo_update.npz"))
    g = {k[len(tag) + 1:]: torch.from_numpy(np.asarray(z[k])) for k in z.files if k.startswith(tag + "_")}
    init = {k[len("init__"):]: v for k, v in g.items() if k.startswith("init__")}
    final = {k[len("final__"):]: v for k, v in g.items() if k.startswith("final__")}
    buf = {k[len("buf__"):]: v for k, v in g.items() if k.startswith("buf__")}
    log = {k[len("log__"):]: float(v) for k, v in g.items() if k.startswith("log__")}
    return g, init, final, buf, log


def _cfg(use_lip):
    from taco_b200.ppo import PPOConfig
    return PPOConfig(clip=0.2, target_kl=0.5, max_grad=0.5, epochs=40, train_iters=2, lr=1e-3, pi_coef=1.0, vf_coef=0.5, ent_coef=0.01,
                     lr_ratio=0.3, lr_lp_index=0.7, lr_epoch_index=30, use_lipschitz=use_lip, lipschitz_para=2.0,
                     lip_ratio=[1.0, 0.3], lip_lp_index=[0.3, 0.7], lip_epoch_index=[5, 30], diff_value=[0.1, 1.0],
                     diff_lp_index=[0.3, 0.7], diff_epoch_index=[5, 30])


def _run(golden_dir, tag, device, project):
    import types
    from taco_b200.ppo import TorchActorCritic, make_optimizer, ppo_update
    g, init, final, buf, log = _load(golden_dir, tag)
    agent = TorchActorCritic(26, 4, [32, 24], 26, 16, [20])
    assert set(agent.state_dict().keys()) == set(init.keys())                 # same parameter names as PPO_ActorCritic
    agent.load_state_dict(init)
    agent.to(device)
    cfg = _cfg(tag == "lip")
    opt = make_optimizer(agent, cfg)
    b = types.SimpleNamespace(**{k: v.to(device) for k, v in buf.items()})
    env = types.SimpleNamespace(difficulty=0.0)
    idx = [row.tolist() for row in g["idx"]]
    out = ppo_update(agent, opt, b, cfg, int(g["epoch"]), env=env, batch_idx=idx, project=project)
    assert out["optim_steps"] == int(g["optim_step"]) and not out["early_stop"]
    assert abs(env.difficulty - float(g["difficulty"])) < 1e-6
    for name in ("policy_gradient_loss", "value_loss", "entropy_loss", "sum_loss", "approx_kl", "learning_rate", "lipschitz_para", "difficulty", "mean_value", "explained_variance"):
        assert out[name] == pytest.approx(log[name], rel=2e-4, abs=2e-6), name
    worst = 0.0
    for k, v in agent.state_dict().items():
        worst = max(worst, (v.cpu() - final[k]).abs().max().item())
    assert worst <= 2e-5, worst
    moved = max((final[k] - init[k]).abs().max().item() for k in init)
    assert moved > 1e-3                                                         # the update did something


def test_update_matches_reference_cpu(golden_dir):
    _run(golden_dir, "plain", "cpu", None)


def test_update_with_projection_matches_reference_cpu(golden_dir):
    """The projection injected here is the oracle's restatement of PPO.spectral_normalize_actors (torch SVD, CPU)."""
    from oracle import actor as oa

    def project(params, c):
        for p in params:
            if p.dim() >= 2:
                w, _ = oa.spectral_normalize([p.data], c)
                p.data.copy_(w[0])
    _run(golden_dir, "lip", "cpu", project)


def test_schedules_follow_the_reference_formulas():
    from taco_b200.ppo import PPOConfig, schedules
    cfg = PPOConfig()                                                           # reference defaults (ppo_asymmetry.py:26-33)
    lr, lip, diff = schedules(cfg, 0)
    assert lr == pytest.approx(3e-4) and lip == pytest.approx(5.0) and diff == pytest.approx(0.1)
    lr, lip, diff = schedules(cfg, 499)
    assert lr == pytest.approx(0.3 * 3e-4) and lip == pytest.approx(0.3 * 5.0) and diff == pytest.approx(1.0)
    lr, lip, diff = schedules(cfg, 250)                                         # learning_process 0.5: both ramps half way
    assert lr == pytest.approx(min((0.3 - 1) / 0.7 * 0.5 + 1, (0.3 - 1) / 350 * 250 + 1) * 3e-4)
    assert lip == pytest.approx(min(0.65, (0.3 - 1.0) / 400 * 150 + 1.0) * 5.0)
    assert diff == pytest.approx(max(0.55, 0.9 / 400 * 150 + 0.1))


@pytest.mark.gpu
def test_update_with_device_projection_matches_reference_gpu(golden_dir):
    _run(golden_dir, "lip", "cuda", None)


@pytest.mark.gpu
def test_trained_weights_reach_the_rollout_kernels():
    """sync_rollout_nets: the rollout kernels then compute what the torch modules compute (FP32 paths, <= 5e-6 against the modules
    evaluated on the CPU: cuDNN's fused LSTM cell deviates by ~5e-5 from the reference arithmetic, our kernel does not)."""
    from taco_b200 import ActorMLP, CriticLSTM
    from taco_b200.ppo import TorchActorCritic, sync_rollout_nets
    torch.manual_seed(3)
    agent = TorchActorCritic(26, 4, [64, 64], 26, 32, [64]).cuda()
    with torch.no_grad():
        for p in agent.parameters():
            p.add_(0.05 * torch.randn_like(p))
    actor, critic = ActorMLP(26, [64, 64], 4), CriticLSTM(26, 5, 32, [64])
    sync_rollout_nets(agent, actor, critic)
    obs, states = torch.randn(300, 1, 26), torch.randn(300, 5, 26)
    with torch.no_grad():
        _, _, value, mean, _ = agent.cpu().evaluate(obs, states, torch.zeros(300, 4))
    assert (actor.forward(obs.cuda()).cpu() - mean).abs().max().item() <= 5e-6
    assert (critic.forward(states.cuda()).cpu() - value).abs().max().item() <= 5e-6
    actor.close(); critic.close()


# ---------------------------------------------------------------------------------------------- world_size 2 (gloo, CPU)
def _ddp_worker(rank, world, port, golden_dir, out_dir):
    import types
    import torch.distributed as dist
    from taco_b200.ppo import TorchActorCritic, make_optimizer, ppo_update
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    g, init, final, buf, log = _load(golden_dir, "plain")
    agent = TorchActorCritic(26, 4, [32, 24], 26, 16, [20])
    agent.load_state_dict(init)
    cfg = _cfg(False)
    b = types.SimpleNamespace(**buf)
    # each rank takes its half of every minibatch (an env-sharded job: every rank holds its own envs' samples)
    idx = [row.tolist()[rank::world] for row in g["idx"]]
    out = ppo_update(agent, make_optimizer(agent, cfg), b, cfg, int(g["epoch"]), batch_idx=idx)
    torch.save({"state": agent.state_dict(), "out": out}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_two_rank_update_equals_single_process_update(golden_dir, tmp_path):
    """Gradients averaged over two ranks that each hold half of every minibatch == the reference's single-process update on
    the whole minibatch (all losses are means over equally many samples); both ranks end with identical parameters."""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_ddp_worker, args=(2, port, golden_dir, str(tmp_path)), nprocs=2, join=True)
    _, init, final, _, log = _load(golden_dir, "plain")
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    for k in final:
        assert torch.equal(r0["state"][k], r1["state"][k]), k
        assert (r0["state"][k] - final[k]).abs().max().item() <= 2e-5, k
    assert r0["out"]["optim_steps"] == 6
    assert r0["out"]["sum_loss"] == pytest.approx(log["sum_loss"], rel=2e-4, abs=2e-6)

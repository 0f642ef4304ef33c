"""Zero-copy rollout storage (taco_env_attach_rollout) and the host-sync-free collection loop (collect_rollout).  GPU only.

The ring path must store, bit for bit, what the reference loop stores by copying after every step
(ppo_asymmetry.py:305-342 -> buffer_asymmetry.py:49-68): obs / states seen BEFORE step s in slot s, reward / done / time-out of
step s in row s -- also across a rewind and for env counts that are not a multiple of the CTA size."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _classic_rollout(env, H, t0):
    obs, states, rew, done, tout = [], [], [], [], []
    for s in range(H):
        obs.append(env.obs_buf.clone()); states.append(env.states_buf.clone())
        _, r, d, ex = env.step(env.random_actions(t0 + s))
        rew.append(r.clone()); done.append(d.clone()); tout.append(ex["time_outs"].clone())
    return torch.stack(obs), torch.stack(states), torch.stack(rew), torch.stack(done), torch.stack(tout)


@pytest.mark.parametrize("task,n", [("flip", 1000), ("mix", 4096)])
def test_ring_stores_what_the_copy_loop_stores(task, n):
    from taco_b200 import FpvVecTask, RolloutBuffer, make_cfg
    H = 7
    cfg = make_cfg(task, n, **{"env.maxEpisodeLength": 12})          # short episodes: time-outs and resets inside the window
    a = FpvVecTask(cfg, seed=5)
    b = FpvVecTask(cfg, seed=5)
    buf = RolloutBuffer(n, 26, 1, 26, 5, 4, H, 1, 0.99, 0.95, "cuda:0")
    for t in range(3):                                               # some history before attaching
        a.step(a.random_actions(t)); b.step(b.random_actions(t))
    b.attach_rollout(buf)
    n_done = 0
    for rollout in range(3):
        t0 = 3 + rollout * H
        if rollout:
            b.rewind_rollout()
        o, st, r, d, to = _classic_rollout(a, H, t0)
        for s in range(H):
            assert torch.equal(b.obs_buf, buf.obs_ring[s])
            b.step(b.random_actions(t0 + s))
        assert torch.equal(buf.obs_buf, o) and torch.equal(buf.states_buf, st)
        assert torch.equal(buf.rew_buf.squeeze(-1), r)
        assert torch.equal(buf.done_buf.squeeze(-1), d.float())
        assert torch.equal(buf.timeout_buf.squeeze(-1).bool(), to)
        assert torch.equal(b.obs_buf, a.obs_buf) and torch.equal(b.states_buf, a.states_buf)
        n_done += int(d.sum())
    assert n_done > 0
    with pytest.raises(RuntimeError, match="ring is full"):
        b.step(b.random_actions(0))
    b.detach_rollout()
    assert torch.equal(b.obs_buf, a.obs_buf) and torch.equal(b.states_buf, a.states_buf)
    a.step(a.random_actions(99)); b.step(b.random_actions(99))
    assert torch.equal(b.states_buf, a.states_buf)
    assert np.array_equal(a.export_state(), b.export_state())
    a.close(); b.close()


def test_collect_rollout_matches_reference_style_loop():
    """collect_rollout == act -> clip -> step -> store copies -> bootstrap -> GAE -> normalise, checked with the oracle."""
    from taco_b200 import ActorMLP, FpvVecTask, RolloutBuffer, collect_rollout, make_cfg
    from oracle import gae as og
    n, H, gamma, lam = 2048, 8, 0.99, 0.95
    cfg = make_cfg("mix", n, **{"env.maxEpisodeLength": 6})
    env = FpvVecTask(cfg, seed=11)
    gen = torch.Generator().manual_seed(0)
    sizes = [26, 64, 64, 4]
    ws = [torch.randn(sizes[i + 1], sizes[i], generator=gen) * 0.2 for i in range(3)]
    bs = [torch.zeros(sizes[i + 1]) for i in range(3)]
    actor = ActorMLP(26, [64, 64], 4)
    actor.load(ws, bs, log_std=torch.full((4,), -0.5))
    critic = torch.nn.Linear(5 * 26, 1).cuda()
    value_fn = lambda obs, states: critic(states.reshape(states.size(0), -1)).detach()
    buf = RolloutBuffer(n, 26, 1, 26, 5, 4, H, 1, gamma, lam, "cuda:0")
    for rollout in range(2):
        stats = collect_rollout(env, actor, buf, value_fn, seed=3, tensor_cores=True)
        assert buf.step == H and stats[7].item() == n * H
        # the stored transitions are self-consistent: obs of slot s+1 is what stepping slot s with act_buf[s] produced
        assert torch.equal(env.obs_buf, buf.obs_ring[H])
        assert torch.all(buf.act_buf.isfinite()) and torch.all(buf.logp_buf.isfinite())
        val = torch.stack([value_fn(buf.obs_ring[s], buf.states_ring[s]) for s in range(H)]).cpu()
        assert torch.equal(buf.value_buf.cpu(), val)
        last = value_fn(buf.obs_ring[H], buf.states_ring[H]).cpu()
        rew, done, tout = buf.rew_buf.cpu().squeeze(-1), buf.done_buf.cpu().squeeze(-1), buf.timeout_buf.cpu().squeeze(-1)
        aug = og.bootstrap_timeouts(rew, val.squeeze(-1), done, tout, gamma)
        adv, ret = og.gae(aug.unsqueeze(-1), done.unsqueeze(-1), val, last, gamma, lam)
        assert torch.equal(buf.ret_buf.cpu(), ret)
        torch.testing.assert_close(buf.adv_buf.cpu(), og.normalize(adv), rtol=0, atol=5e-6)
        assert (tout.bool() & (done != 0)).any(), "window must contain truncated episodes"
        assert stats[1].item() == done.sum().item() and stats[2].item() == (tout.bool() & (done != 0)).sum().item()
    env.close(); actor.close()


@pytest.mark.parametrize("tensor_cores", [False, True])
def test_collect_rollout_with_native_critic(tensor_cores):
    """agent.act of the reference (actor sample + critic value, nets_asymmetry.py:326-352) entirely in the library: the critic
    kernels write value_buf rows in place; values equal a separate CriticLSTM.forward on the stored state histories."""
    from taco_b200 import ActorMLP, CriticLSTM, FpvVecTask, RolloutBuffer, collect_rollout, make_cfg
    from oracle import critic as oc, gae as og
    n, H, gamma, lam = 1500, 6, 0.99, 0.95
    env = FpvVecTask(make_cfg("flip", n), seed=5)
    gen = torch.Generator().manual_seed(1)
    sizes = [26, 64, 64, 4]
    actor = ActorMLP(26, [64, 64], 4)
    actor.load([torch.randn(sizes[i + 1], sizes[i], generator=gen) * 0.2 for i in range(3)], [torch.zeros(sizes[i + 1]) for i in range(3)])
    hid, mlp = 32, [64, 64]
    lstm = [(torch.randn(4 * hid, 26, generator=gen) * 0.3, torch.randn(4 * hid, hid, generator=gen) * 0.2,
             torch.randn(4 * hid, generator=gen) * 0.1, torch.randn(4 * hid, generator=gen) * 0.1)]
    cs = [hid] + mlp + [1]
    cw = [torch.randn(cs[i + 1], cs[i], generator=gen) / cs[i] ** 0.5 for i in range(3)]
    cb = [torch.randn(cs[i + 1], generator=gen) * 0.1 for i in range(3)]
    critic = CriticLSTM(26, 5, hid, mlp)
    critic.load(lstm, cw, cb)
    buf = RolloutBuffer(n, 26, 1, 26, 5, 4, H, 1, gamma, lam, "cuda:0")
    collect_rollout(env, actor, buf, critic, seed=3, tensor_cores=tensor_cores)
    val = torch.stack([critic.forward(buf.states_ring[s], tensor_cores=tensor_cores) for s in range(H)])
    assert torch.equal(buf.value_buf, val)
    ref = torch.stack([oc.critic_forward(buf.states_ring[s].cpu(), lstm, cw, cb) for s in range(H)])
    tol = 3e-2 if tensor_cores else 5e-6
    assert (val.cpu() - ref).abs().max().item() <= tol
    last = critic.forward(buf.states_ring[H], tensor_cores=tensor_cores).cpu()
    rew, done, tout = buf.rew_buf.cpu().squeeze(-1), buf.done_buf.cpu().squeeze(-1), buf.timeout_buf.cpu().squeeze(-1)
    aug = og.bootstrap_timeouts(rew, val.cpu().squeeze(-1), done, tout, gamma)
    adv, ret = og.gae(aug.unsqueeze(-1), done.unsqueeze(-1), val.cpu(), last, gamma, lam)
    assert torch.equal(buf.ret_buf.cpu(), ret)
    env.close(); actor.close(); critic.close()


@pytest.mark.parametrize("task,dr", [("flip", False), ("mix", True)])
def test_graphed_rollout_replays_are_bit_identical_to_the_eager_loop(task, dr):
    """GraphedRollout: the whole rollout as one CUDA-graph replay, the step index in a device counter.  Three rollouts (one eager
    in the constructor + two replays; weights, log_std and the env difficulty changed in between) equal three eager collect_rollout
    calls bit for bit."""
    from taco_b200 import ActorMLP, CriticLSTM, FpvVecTask, GraphedRollout, RolloutBuffer, collect_rollout, make_cfg
    n, H = 3000, 7
    gen = torch.Generator().manual_seed(2)
    sizes = [26, 64, 64, 4]
    aw = [[torch.randn(sizes[i + 1], sizes[i], generator=gen) * 0.25 for i in range(3)] for _ in range(2)]
    ab = [torch.zeros(sizes[i + 1]) for i in range(3)]
    hid, mlp = 32, [64]
    lstm = [(torch.randn(4 * hid, 26, generator=gen) * 0.3, torch.randn(4 * hid, hid, generator=gen) * 0.2,
             torch.randn(4 * hid, generator=gen) * 0.1, torch.randn(4 * hid, generator=gen) * 0.1)]
    cs = [hid] + mlp + [1]
    cw = [torch.randn(cs[i + 1], cs[i], generator=gen) / cs[i] ** 0.5 for i in range(2)]
    cb = [torch.randn(cs[i + 1], generator=gen) * 0.1 for i in range(2)]
    runs = {}
    for mode in ("eager", "graph"):
        env = FpvVecTask(make_cfg(task, n, domain_randomization=dr, **{"env.maxEpisodeLength": 9}), seed=21)
        actor, critic = ActorMLP(26, [64, 64], 4), CriticLSTM(26, 5, hid, mlp)
        actor.load(aw[0], ab, log_std=torch.full((4,), -0.3)); critic.load(lstm, cw, cb)
        buf = RolloutBuffer(n, 26, 1, 26, 5, 4, H, 1, 0.99, 0.95, "cuda:0")
        snaps, gr = [], None
        for k in range(3):
            if k == 2:
                # new weights, a new log_std (a trained nn.Parameter in the reference, nets_asymmetry.py:315) and a new difficulty
                # (written by the trainer every epoch, ppo_asymmetry.py:173-175) all reach the replayed kernels
                actor.load(aw[1], ab, log_std=torch.full((4,), -0.55))
                env.difficulty = 0.35
            if mode == "eager":
                stats = collect_rollout(env, actor, buf, critic, seed=4, tensor_cores=True)
            elif k == 0:
                gr = GraphedRollout(env, actor, buf, critic, seed=4, tensor_cores=True)
                stats = gr.first_stats
            else:
                stats = gr.run()
            torch.cuda.synchronize()
            snaps.append([t.clone() for t in (buf.obs_ring, buf.states_ring, buf.act_buf, buf.logp_buf, buf.value_buf, buf.rew_buf, buf.done_buf,
                                              buf.timeout_buf, buf.ret_buf, buf.adv_buf, stats)])
        if gr is not None:
            gr.close()
            assert env.step_count == 3 * H and env.step_counter() == (0, 3 * H)
        env.rewind_rollout()
        o, r, x, e = env.step(env.random_actions(0))                               # eager stepping continues after the graph
        snaps.append([o["states"].clone(), r.clone(), x.clone()])
        runs[mode] = snaps
        env.close(); actor.close(); critic.close()
    for k, (a, b) in enumerate(zip(runs["eager"], runs["graph"])):
        for j, (ta, tb) in enumerate(zip(a, b)):
            assert torch.equal(ta, tb), (k, j)
    assert runs["eager"][2][6].sum() > 0 and runs["eager"][2][7].sum() > 0                 # the window holds resets and time-outs

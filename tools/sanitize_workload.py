"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every hand-written kernel
on shapes that exercise its edge paths.  usage: compute-sanitizer --tool <tool> python tools/sanitize_workload.py [part]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taco_b200  # noqa: E402
from taco_b200 import ActorMLP, CriticLSTM, RolloutBuffer, make_cfg  # noqa: E402

part = sys.argv[1] if len(sys.argv) > 1 else "all"
gen = torch.Generator().manual_seed(0)


def step_part():
    # ragged env counts (not a multiple of the 128-env CTA), every DR / noise switch, all four tasks
    for task, n, kw in (("mix", 1000, dict(domain_randomization=True, observation_noise=True, rotor_noise=True)),
                        ("flip", 129, {}), ("rotate", 1, {}), ("pos", 4096 + 77, {})):
        env = taco_b200.FpvVecTask(make_cfg(task, n, **kw), "cuda:0", "cuda:0", -1, True, seed=3)
        env.reset()
        for t in range(3):
            env.step(env.random_actions(t))
        env.stats()
        torch.cuda.synchronize()
        env.close()
    # natural, sparse resets (episodes end at different steps -> one or two resetting lanes per warp: the warp-generated reset draws
    # through the shared-memory mailbox) in both launch shapes: <= 148 CTAs (copy warps) and a larger grid (copy after the step)
    for n, steps in ((4096, 70), (148 * 128 + 999, 45)):
        env = taco_b200.FpvVecTask(make_cfg("flip", n), "cuda:0", "cuda:0", -1, True, seed=5)
        for t in range(steps):
            env.step(env.random_actions(t))
        st = env.stats().cpu().tolist()
        assert st[1] > 0, "no episode ended: the sparse-reset path was not exercised"
        env.close()
    print("step ok")


def nets_part():
    sizes = [26, 256, 256, 256, 4]
    ws = [torch.randn(sizes[i + 1], sizes[i], generator=gen) / sizes[i] ** 0.5 for i in range(4)]
    bs = [torch.randn(sizes[i + 1], generator=gen) * 0.1 for i in range(4)]
    actor = ActorMLP(26, sizes[1:-1], 4)
    actor.load(ws, bs, lipschitz_const=4.0)
    hid, mlp = 64, [256, 256, 256]
    lstm = [(torch.randn(4 * hid, 26, generator=gen) * 0.3, torch.randn(4 * hid, hid, generator=gen) * 0.2,
             torch.randn(4 * hid, generator=gen) * 0.1, torch.randn(4 * hid, generator=gen) * 0.1)]
    cs = [hid] + mlp + [1]
    cw = [torch.randn(cs[i + 1], cs[i], generator=gen) / cs[i] ** 0.5 for i in range(4)]
    cb = [torch.randn(cs[i + 1], generator=gen) * 0.1 for i in range(4)]
    critic = CriticLSTM(26, 5, hid, mlp)
    critic.load(lstm, cw, cb)
    # 1000 envs: one tile per CTA, partial last tile; 148*2*128 + 77: a pair of tiles per CTA, partial last tile
    for n in (1000, 148 * 2 * 128 + 77):
        obs = torch.randn(n, 1, 26, generator=gen).cuda()
        states = torch.randn(n, 5, 26, generator=gen).cuda()
        for tc in (True, False):
            if not tc and n > 2000:
                continue
            actor.forward(obs, tensor_cores=tc)
            actor.act(obs, step_index=1, tensor_cores=tc)
            critic.forward(states, tensor_cores=tc)
        torch.cuda.synchronize()
    actor.close(); critic.close()
    print("nets ok")


def gae_part():
    H, N = 6, 500
    buf = RolloutBuffer(N, 26, 1, 26, 5, 4, H, 1, 0.99, 0.95, "cuda:0")
    buf.rew_buf.copy_(torch.rand(H, N, 1, generator=gen) * 0.02)
    buf.value_buf.copy_(torch.randn(H, N, 1, generator=gen) * 0.3)
    buf.done_buf.copy_((torch.rand(H, N, 1, generator=gen) < 0.1).float())
    buf.compute_returns_and_advantage((torch.randn(N, 1, generator=gen) * 0.3).cuda())
    torch.cuda.synchronize()
    print("gae ok")


def ppo_part():
    # the native PPO update: TMA-fed tcgen05 GEMM (all epilogues, split-K), gather, loss, LSTM backward, reductions, Adam, projection
    import types
    from taco_b200.ppo import PPOConfig, TorchActorCritic
    from taco_b200.ppo_native import NativePPO, gemm_selftest
    for (m, n, k, sp) in ((300, 48, 96, 1), (256, 32, 2048, 7), (4, 256, 1024, 5)):
        a = (torch.randn(m, k, device="cuda") * 0.5).bfloat16()
        b = (torch.randn(n, k, device="cuda") * 0.5).bfloat16()
        gemm_selftest(a, b, sp)
    for (m, n, k, sp) in ((256, 64, 1024, 3), (48, 256, 640, 2)):          # transposed (MN-major) operands; the first shape stores through TMA
        at = (torch.randn(k, m, device="cuda") * 0.5).bfloat16()
        bt = (torch.randn(k, n, device="cuda") * 0.5).bfloat16()
        gemm_selftest(at, bt, sp, transposed=True)
    agent = TorchActorCritic(26, 4, [64, 48], 26, 64, [32]).cuda()
    H, N = 4, 128
    g = torch.Generator(device="cuda").manual_seed(0)
    rnd = lambda *s: torch.randn(*s, device="cuda", generator=g)
    buf = types.SimpleNamespace(obs_buf=rnd(H, N, 1, 26), states_buf=rnd(H, N, 5, 26), act_buf=rnd(H, N, 4).clamp(-1, 1), value_buf=rnd(H, N, 1),
                                ret_buf=rnd(H, N, 1), adv_buf=rnd(H, N, 1), logp_buf=rnd(H, N, 1) * 0.1 - 4.0)
    nat = NativePPO(agent, 256)
    perm = torch.randperm(H * N, device="cuda").view(2, -1)
    nat.update(buf, PPOConfig(train_iters=2, use_lipschitz=True, lipschitz_para=1.0, target_kl=1e9), 0, batch_idx=[perm[0], perm[1]])
    nat.update(buf, PPOConfig(train_iters=2, target_kl=1e-9), 1, batch_idx=[perm[0], perm[1]])        # early stop: the no-op launches
    torch.cuda.synchronize()
    nat.close()
    print("ppo ok")


if part in ("all", "step"):
    step_part()
if part in ("all", "nets"):
    nets_part()
if part in ("all", "gae"):
    gae_part()
if part in ("all", "ppo"):
    ppo_part()

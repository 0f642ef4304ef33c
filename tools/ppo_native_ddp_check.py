"""NativePPO under torch.distributed (NCCL): every rank updates on its share of every minibatch; the loss sums and the flat gradient
are all-reduced.  Checks: all ranks end with bit-identical parameters; they agree with a single-process update on the whole
minibatches (same arithmetic, different summation order); the early-stop decision is common.
usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/ppo_native_ddp_check.py"""
import json
import os
import sys
import types

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from taco_b200.ppo import PPOConfig, TorchActorCritic  # noqa: E402
from taco_b200.ppo_native import NativePPO  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
dev = f"cuda:{local}"
torch.manual_seed(5)
agent = TorchActorCritic(26, 4, [128, 64], 26, 64, [64]).to(dev)
H, N, mb = 16, 512, 2
gen = torch.Generator(device=dev).manual_seed(7)                 # the same data on every rank; each rank USES its share
rnd = lambda *s: torch.randn(*s, device=dev, generator=gen)
buf = types.SimpleNamespace(obs_buf=rnd(H, N, 1, 26) * 0.6, states_buf=rnd(H, N, 5, 26) * 0.6, act_buf=rnd(H, N, 4).clamp(-1.5, 1.5),
                            value_buf=torch.zeros(H, N, 1, device=dev), ret_buf=rnd(H, N, 1) * 0.4, adv_buf=rnd(H, N, 1))
with torch.no_grad():
    buf.logp_buf = (agent.evaluate(buf.obs_buf.view(-1, 1, 26), buf.states_buf.view(-1, 5, 26), buf.act_buf.view(-1, 4))[0] + 0.05 * rnd(H * N)).view(H, N, 1)
perm = torch.randperm(H * N, device=dev, generator=gen).view(mb, -1)
cfg = PPOConfig(train_iters=2, lr=1e-3, use_lipschitz=True, lipschitz_para=1.5, target_kl=1e9, ent_coef=0.01)
share = [perm[i][rank::world].contiguous() for i in range(mb)]
nat = NativePPO(agent, share[0].numel(), device=dev)
out = nat.update(buf, cfg, 3, batch_idx=share)
mine = nat.params.clone()
gathered = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(gathered, mine)
res = {"world": world, "ranks_bit_identical": all(torch.equal(g, gathered[0]) for g in gathered), "optim_steps": out["optim_steps"], "kl": out["approx_kl"]}
if rank == 0:
    single = NativePPO(agent, perm.shape[1], device=dev)
    o1 = single.update(buf, cfg, 3, batch_idx=[perm[i] for i in range(mb)], group=False)
    init = torch.cat([p.detach().flatten() for n, p in agent.named_parameters()])      # order differs from the flat vector: compare deltas by name
    vs, vm = single._views(single.params), nat._views(mine)
    sd = dict(agent.named_parameters())
    d_s = torch.cat([(vs[n] - sd[n].detach()).flatten() for n in vs])
    d_m = torch.cat([(vm[n] - sd[n].detach()).flatten() for n in vs])
    res.update({"cosine_vs_single_process": float(torch.nn.functional.cosine_similarity(d_s, d_m, dim=0)),
                "rel_l2_vs_single_process": float((d_s - d_m).norm() / d_s.norm()), "single_kl": o1["approx_kl"],
                "losses": [out["policy_gradient_loss"], o1["policy_gradient_loss"], out["value_loss"], o1["value_loss"]]})
    print(json.dumps(res))
    single.close()
nat.close()
dist.destroy_process_group()

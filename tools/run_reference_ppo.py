#!/usr/bin/env python
"""Drop-in proof with the reference's own consumer: the UNMODIFIED ``PPO`` class of IsaacGymEnvs/algorithms/ppo_asymmetry.py
(its rollout loop ``PPO.run``, :286-393, its ``update``, its replay buffer, its ``PPO_ActorCritic``) trains through
``taco_b200.FpvVecTask`` -- the one edit a maintainer makes is the ``isaacgym_task_map`` line shown in INTEGRATION.md.

    bash tools/install_reference.sh                      (build container: copies algorithms/ into the git-ignored baseline/_ref/)
    python tools/run_reference_ppo.py --task flip --epochs 20 --out profiles/reference_ppo_flip_r02.jsonl      (GPU box)

The actor / critic / PPO dictionaries are built the way train_fpv_asymmetry_ppo.py:376-538 builds them from the (missing) YAML:
MLP actor without encoder, LSTM critic encoder (README.md:60-66); sizes and PPO hyper-parameters are our stated choices.
The reference ends ``run`` by re-loading its whole-module pickles ``model_0.pt`` / ``model_1.pt`` with ``torch.load`` and tracing
``actor_{0,1}.pt`` (:385-393); torch >= 2.6 needs TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD=1 for that, which is set here (environment,
not a code change).  Per-epoch numbers are taken from the reference's own TensorBoard calls by wrapping its SummaryWriter.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD", "1")


def find_reference():
    for cand in (os.path.join(ROOT, "baseline", "_ref"), os.path.join(os.environ.get("TACO_REFERENCE", "/root/reference"), "IsaacGymEnvs")):
        if os.path.exists(os.path.join(cand, "algorithms", "ppo_asymmetry.py")):
            return cand
    return None


def build(task, num_envs, epochs, horizon, log_dir, seed=42, actor_hidden=(256, 256, 256), critic_hidden=(256, 256, 256), lstm_hidden=64,
          mini_batch_num=4, train_iters=4, use_lipschitz=True):
    """(env, reference PPO object).  Imports the reference package from baseline/_ref (or the reference tree)."""
    ref = find_reference()
    if ref is None:
        raise RuntimeError("reference algorithms/ not found: run tools/install_reference.sh in the build container")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import torch.nn as nn
    import taco_b200
    from algorithms.ppo_asymmetry import PPO                       # the reference's own classes, unmodified
    from algorithms.nets_asymmetry import PPO_ActorCritic
    cfg = taco_b200.make_cfg(task, num_envs)
    env = taco_b200.isaacgym_task_map["Fpv_" + task](cfg, "cuda:0", "cuda:0", -1, True, False, False)     # train_fpv_asymmetry_ppo.py:363-371
    para = {
        "actor_critic_mlp_dict": {"actor_hidden_sizes": list(actor_hidden), "critic_hidden_sizes": list(critic_hidden),
                                  "actor_input_dim": env.num_obs * env.len_obs, "actor_output_dim": env.num_acts,
                                  "critic_input_dim": env.num_states * env.len_states, "critic_output_dim": 1, "activation": nn.ReLU},
        "use_actor_encoder": False, "use_critic_encoder": True, "share_encoder": False,
        "actor_encoder_type": "LSTM", "critic_encoder_type": "LSTM",
        "critic_encoder_dict": {"encoder_type": "LSTM", "input_size": env.num_states, "output_size": lstm_hidden, "num_layers": 1,
                                "bidirectional": False},
    }
    ppo = PPO(env=env, actor_critic=PPO_ActorCritic, actor_critic_para_dict=para, epochs=epochs, horizon_len=horizon,
              train_iters=train_iters, mini_batch_num=mini_batch_num, seed=seed, use_lipschitz=use_lipschitz, lipschitz_para=4,
              difficulty_schedule=False, diff_value=[1.0, 1.0], log_dir=log_dir, log_interval=max(epochs // 4, 1), device="cuda:0")
    return env, ppo


class RecordingWriter:
    """Wraps the reference's SummaryWriter: every add_scalar also lands in a per-epoch dict."""

    def __init__(self, inner):
        self.inner, self.rows = inner, {}

    def add_scalar(self, tag, value, step):
        self.rows.setdefault(int(step), {})[tag.rstrip(":")] = float(value)
        self.inner.add_scalar(tag, value, step)

    def __getattr__(self, name):
        return getattr(self.inner, name)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--task", default="flip")
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--epochs", type=int, default=20)
    ap.add_argument("--horizon", type=int, default=64)
    ap.add_argument("--out", default="")
    ap.add_argument("--log-dir", default="/tmp/taco_reference_ppo")
    a = ap.parse_args()
    import torch
    env, ppo = build(a.task, a.envs, a.epochs, a.horizon, a.log_dir)
    ppo.writer = RecordingWriter(ppo.writer)
    t0 = time.time()
    ppo.run()                                                        # ppo_asymmetry.py:286-393, untouched
    torch.cuda.synchronize()
    wall = time.time() - t0
    rows = []
    for ep in sorted(ppo.writer.rows):
        rows.append({"epoch": ep, **ppo.writer.rows[ep]})
    nn_dir = os.path.join(a.log_dir, "nn")
    summary = {"summary": True, "task": a.task, "envs": a.envs, "epochs": a.epochs, "horizon": a.horizon, "wall_s": wall,
               "env_steps": a.envs * a.horizon * a.epochs, "loop_env_steps_per_s": a.envs * a.horizon * a.epochs / wall,
               "optim_steps": int(ppo.optim_step), "files_written": sorted(os.listdir(nn_dir)) if os.path.isdir(nn_dir) else [],
               "consumer": "reference PPO.run (algorithms/ppo_asymmetry.py, unmodified) on taco_b200.FpvVecTask"}
    text = "\n".join(json.dumps(r) for r in rows + [summary])
    if a.out:
        with open(a.out, "w") as fh:
            fh.write(text + "\n")
    print(text)
    env.close()


if __name__ == "__main__":
    main()

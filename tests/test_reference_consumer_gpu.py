"""The reference's OWN consumer drives the drop-in: the unmodified ``PPO`` class of IsaacGymEnvs/algorithms/ppo_asymmetry.py
(rollout loop ``PPO.run`` :286-393, ``update``, ``PPOReplayBuffer``, ``PPO_ActorCritic``) trains through ``FpvVecTask``.
GPU only; needs the git-ignored copy ``baseline/_ref/algorithms`` made by tools/install_reference.sh (it ships with gpurun)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_reset_returns_storage_the_env_never_overwrites():
    """PPO.run keeps reset()'s tensors as its own observation storage and reads them AFTER env.step
    (ppo_asymmetry.py:297-299,326-329); the reference returns fresh clamp() results (vec_task_asymmetry.py:358-359)."""
    import taco_b200
    env = taco_b200.FpvFlip(taco_b200.make_cfg("flip", 256), "cuda:0", "cuda:0", -1, True, False, False)
    d = env.reset()
    obs, states = d["obs"], d["states"]
    for t in range(3):
        env.step(env.random_actions(t))
    assert float(obs.abs().max()) == 0.0 and float(states.abs().max()) == 0.0
    env.close()


def test_reference_ppo_run_trains_through_fpv_vec_task(tmp_path):
    import run_reference_ppo as rr
    if rr.find_reference() is None:
        pytest.skip("baseline/_ref/algorithms absent (tools/install_reference.sh, build container only)")
    env, ppo = rr.build("flip", 512, epochs=3, horizon=8, log_dir=str(tmp_path), actor_hidden=(64, 64), critic_hidden=(64,), lstm_hidden=32,
                        mini_batch_num=2, train_iters=2)
    assert type(ppo).__module__ == "algorithms.ppo_asymmetry" and type(ppo.replay_buffer).__module__ == "algorithms.buffer_asymmetry"
    before = [p.detach().clone() for p in ppo.agent.parameters()]
    ppo.writer = rr.RecordingWriter(ppo.writer)
    ppo.run()
    assert ppo.optim_step > 0 and any(not torch.equal(a, b.detach()) for a, b in zip(before, ppo.agent.parameters()))
    files = set(os.listdir(os.path.join(str(tmp_path), "nn")))
    assert {"model_0.pt", "model_1.pt", "actor_0.pt", "actor_1.pt"} <= files          # ppo_asymmetry.py:369-393
    rows = ppo.writer.rows
    assert len(rows) == 3 and all("Update/approx_kl" in r and "Interact/Reward" in r for r in rows.values())
    # what the reference's buffer stored is what the env produced: frame s+1 of the observation history equals the newest
    # frame of the states history at the same step for a task without observation noise
    buf = ppo.replay_buffer
    assert torch.isfinite(buf.obs_buf).all() and torch.isfinite(buf.ret_buf).all()
    assert torch.equal(buf.obs_buf[1:, :, -1, :], buf.states_buf[1:, :, -1, :])
    assert not torch.equal(buf.obs_buf[1], buf.obs_buf[2])
    env.close()

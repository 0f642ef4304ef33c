"""Generate tests/golden/glue_<task>_<mode>.npz by RUNNING THE REFERENCE'S OWN ENV CLASSES -- FpvPos / FpvRotate /
FpvFlip / FpvMix of IsaacGymEnvs/isaacgymenvs/tasks/fpv_asymmetry.py, on VecTask of tasks/base/vec_task_asymmetry.py,
unmodified -- on the CPU over oracle/fake_gym.py (simulate = oracle/rigid_body, our integrator specification) with
the random draws fed from the shared Philox slot table (oracle/ref_draws.py).  TEST INFRASTRUCTURE.

    python -m oracle.make_golden_glue          (build container only: needs /root/reference)

Two modes per task:
  det   every ``random_*`` switch off (resets deterministic apart from the two draws the reference always makes,
        fpv_asymmetry.py:792,860), maxEpisodeLength 25 so that time-out resets occur inside the run;
  rand  every reset randomisation, per-env domain randomisation (rotor polynomial / response / aero, random delay and
        deploy time, random voltage), rotor noise and observation noise on; maxEpisodeLength 1000.
At step ``JUMP`` the generator sets ``progress_buf`` to 497 for every env (a plain attribute write, the same surgery
the replay applies) so that the progress == 500 command re-draw (fpv_asymmetry.py:587-603,886-901) is crossed
inside the run.  Actions are the shared Philox action stream (tests/parity_util.oracle_actions).

These files pin the FpvBase glue restated in oracle/fpv_env.py (tests/test_oracle_glue.py) and are replayed through
the CUDA step (tests/test_glue_golden_gpu.py).
"""
import os

import numpy as np
import torch

from . import fake_gym, ref_draws, ref_loader
from . import philox as px

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEED = 0x7AC0
N_ENVS = 18
STEPS = 72
JUMP = 56                 # after this many steps: progress_buf <- 497
FULL_AT = (0, 1, 5, 30, STEPS - 1)      # steps whose complete (N, L, 26) history buffers are stored


def glue_cfg(task, mode, num_envs=N_ENVS):
    """The cfg dict of a fixture (also used by the replaying tests)."""
    from taco_b200.config import make_cfg
    if mode == "det":
        cfg = make_cfg(task, num_envs, random_copter_pos=False, random_copter_quat=False, random_copter_vel=False,
                       random_target_pos=False, random_target_yaw=False, random_voltage=False, random_rotor_speed=False,
                       random_command=False)
        cfg["env"]["maxEpisodeLength"] = 25
    else:
        cfg = make_cfg(task, num_envs, domain_randomization=True, observation_noise=True, rotor_noise=True)
        cfg["difficulty"] = 0.7
    cfg["sim"]["use_gpu_pipeline"] = False          # CPU pipeline (vec_task_asymmetry.py:63-69); ignored by the B200 path
    return cfg


def _pinned(base, feeder, log):
    """Subclass of the reference class that only observes: activates the feeder, counts steps, logs the delay index."""

    class Pinned(base):
        def step(self, actions):
            ref_draws.activate(feeder)
            feeder.env = self
            log["idx"], log["act"] = [], []
            out = super().step(actions)
            feeder.step_index += 1
            return out

        def mid_physics_step(self):
            # the index expression of fpv_asymmetry.py:366, evaluated before the reference evaluates it
            idx = torch.clamp(self.actions_remained_length - 1, max=self.mid_step_count).squeeze(-1).long()
            log["idx"].append(torch.where(idx < 0, idx + 100, idx).clone())
            log["act"].append(self.actions_remained_buffer[torch.arange(self.num_envs), :, idx].clone())
            super().mid_physics_step()

    Pinned.__name__ = "Pinned" + base.__name__
    return Pinned


def actions_at(gid, t):
    return torch.from_numpy(px.u01(px.draw(SEED, gid, t, 0, px.STREAM_ACTIONS)) * np.float32(2.0) - np.float32(1.0))


def run_reference(task, mode, steps=STEPS, save=True):
    fpv = fake_gym.load_env_module()
    ref_draws.install(fpv)
    cfg = glue_cfg(task, mode)
    feeder = ref_draws.Feeder(SEED, N_ENVS)
    ref_draws.activate(feeder)
    log = {}
    base = {"pos": fpv.FpvPos, "rotate": fpv.FpvRotate, "flip": fpv.FpvFlip, "mix": fpv.FpvMix}[task]
    env = fake_gym.make_env(task, cfg, cls=_pinned(base, feeder, log))
    env.reset()
    gid = np.arange(N_ENVS, dtype=np.uint64)
    rec = {k: [] for k in ("actions", "obs", "states", "rew", "reset", "time_outs", "progress", "delay_len", "delay_idx",
                           "delayed_actions", "root", "target", "rotor", "volt", "rpy_cont", "command", "flip_radian",
                           "pid_prev", "battery", "poly", "tau", "aero")}
    full = {}
    for t in range(steps):
        if t == JUMP:
            env.progress_buf[:] = 497
        a = actions_at(gid, t)
        obs, rew, reset, extras = env.step(a.clone())
        rec["actions"].append(a)
        rec["obs"].append(obs["obs"][:, -1].clone())
        rec["states"].append(obs["states"][:, -1].clone())
        rec["rew"].append(rew.clone())
        rec["reset"].append(reset.clone())
        rec["time_outs"].append(extras["time_outs"].clone())
        rec["progress"].append(env.progress_buf.clone())
        rec["delay_len"].append(env.actions_remained_length.squeeze(-1).long().clone())
        rec["delay_idx"].append(torch.stack(log["idx"], dim=1))
        rec["delayed_actions"].append(torch.stack(log["act"], dim=1))
        rec["root"].append(env.root_states.clone())
        rec["target"].append(env.target_states[:, 0:7].clone())
        rec["rotor"].append(env.rotor_speed.clone())
        rec["volt"].append(env.battery_voltage.reshape(-1).clone())
        rec["rpy_cont"].append(env.copter_rpy_continuous.clone())
        rec["command"].append(env.command.clone())
        rec["flip_radian"].append(env.flip_radian.clone() if hasattr(env, "flip_radian") else torch.zeros(N_ENVS))
        rec["pid_prev"].append(env.angvel_controller.previous_error.clone())
        b = env.battery_dynamics
        rec["battery"].append(torch.cat((b.u_1, b.E_c, b.time), dim=1).clone())
        rec["poly"].append(env.rotor_dynamics.omega_para.clone())
        rec["tau"].append(env.rotor_dynamics.response_time.clone())
        a_ = env.aero_dynamics
        rec["aero"].append(torch.cat((a_.para_force_torque, a_.para_d, a_.para_t), dim=1).clone())
        if t in FULL_AT:
            full[f"obs_full_{t}"] = obs["obs"].clone()
            full[f"states_full_{t}"] = obs["states"].clone()
    out = {k: torch.stack(v).numpy() for k, v in rec.items()}
    out.update({k: v.numpy() for k, v in full.items()})
    out["meta"] = np.array([SEED, N_ENVS, STEPS, JUMP], dtype=np.int64)
    n_reset = int(out["reset"].sum())
    n_to = int(out["time_outs"].sum())
    print(f"glue_{task}_{mode}: {feeder.calls} draws fed, {n_reset} resets, {n_to} time-outs, "
          f"finite={bool(np.isfinite(out['states']).all())}")
    if save:
        np.savez_compressed(os.path.join(OUT, f"glue_{task}_{mode}.npz"), **out)
    return out


def main():
    assert ref_loader.available(), "needs the reference tree (build container only)"
    torch.set_num_threads(1)
    for task in ("pos", "rotate", "flip", "mix"):
        for mode in ("det", "rand"):
            run_reference(task, mode)
    for f in sorted(os.listdir(OUT)):
        if f.startswith("glue_"):
            print("  ", f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()

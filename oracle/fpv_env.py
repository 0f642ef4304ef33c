"""RefFpvEnv -- CPU restatement of one ``VecTask.step`` of the TACO fpv_asymmetry tasks
(pos / rotate / flip / mix).  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows, in order:
  VecTask.step                IsaacGymEnvs/isaacgymenvs/tasks/base/vec_task_asymmetry.py:290-334 (VT)
  FpvBase.pre/mid/post step   IsaacGymEnvs/isaacgymenvs/tasks/fpv_asymmetry.py:317-388            (FPV)
  refresh_state               FPV:334-360
  observation/state frames    FPV:390-421 + task overrides :711-714, :766-771, :830-838, :929-946
  resets                      FPV:475-603 + :725-759, :783-821, :850-917, :981-1112
using oracle.dynamics / oracle.rewards / oracle.leaf_math (pinned to the reference's torch
modules by tests/golden) and oracle.rigid_body (our PhysX stand-in).

Pinned against the reference's own FpvPos / FpvRotate / FpvFlip / FpvMix run over oracle/fake_gym.py
(tests/golden/glue_*.npz, tests/test_oracle_glue.py).  Differences from the reference that are deliberate and
documented in DESIGN.md:
  * resets are expressed with boolean masks instead of index lists (same result);
  * random numbers come from the Philox slot table below instead of torch's global RNG (oracle/ref_draws.py maps
    every reference call site onto the table);
  * "reset is visible immediately": the root state written by a reset is what the next
    refresh_state sees (PhysX-internal in the reference, FPV:508 / quirk 17 in SURVEY.md);
  * unless ``reference_exact=True``: the angular velocity is carried in body coordinates across the control
    sub-steps of one RL step, and the random attitude draws use ``sincos_draw`` (both: see ``__init__``).
The dense (N,4,100) delay buffer is kept exactly as in the reference (FPV:189,326-331,
366,378-380) so that the kernel's compressed action queue is checked against it.
"""
import math

import numpy as np
import torch

from . import philox as px
from . import dynamics as dyn
from . import rewards as rw
from . import rigid_body as rb
from .leaf_math import (qmul, qconj, qrot, euler_xyz, quat_from_euler, rotmat9, rand_range, sincos_draw)

TASK_POS, TASK_ROTATE, TASK_FLIP, TASK_MIX = 0, 1, 2, 3
TASK_BY_NAME = {"pos": TASK_POS, "rotate": TASK_ROTATE, "flip": TASK_FLIP, "mix": TASK_MIX}
TWO_PI = 2 * math.pi

# Philox slot table for STREAM_RESET (word index inside the 4x32 block in brackets)
#  0: copter pos x,y,z | flip spin-direction coin          5: rotor poly factors 0..3
#  1: attitude euler a,b,yaw | delay integer draw          6: poly factor 4 | k_f, k_tau, d_x factors
#  2: linvel x,y,z | battery E_c                           7: d_y, k_th factors
#  3: angvel x,y,z | target yaw                            8: rotor response times 0..3
#  4: target pos x,y,z                                     9: initial rotor speeds 0..3
N_RESET_SLOTS = 10


def f32(x):
    return torch.tensor(np.asarray(x, dtype=np.float32))


class RefFpvEnv:
    def __init__(self, cfg, env_offset=0, num_envs_global=None, seed=0, reference_exact=False):
        """``reference_exact=True`` switches off the two places where the specification the CUDA kernel implements
        deliberately departs from what the reference classes compute over a simulator (both documented in DESIGN.md):
        the angular velocity goes through the world-frame root state between control sub-steps (R(q) w_b, then
        R(q)^T of that, fpv_asymmetry.py:350) instead of being carried in body coordinates, and the random attitude
        draws use torch.sin / torch.cos (torch_utils.py:199-213) instead of the polynomial ``sincos_draw``.  In that
        mode this class is held to trajectories recorded from the reference's own FpvPos / FpvRotate / FpvFlip /
        FpvMix (tests/golden/glue_*.npz, tests/test_oracle_glue.py)."""
        e = cfg["env"]
        self.reference_exact = bool(reference_exact)
        self._sincos = None if reference_exact else sincos_draw
        self.cfg = cfg
        self.N = int(e["numEnvs"])
        self.env_offset = int(env_offset)
        self.N_global = int(num_envs_global) if num_envs_global is not None else self.N
        self.gid = np.arange(self.env_offset, self.env_offset + self.N, dtype=np.uint64)
        self.seed = int(seed)
        self.task_mode = TASK_BY_NAME[cfg["task_mode"]]
        self.max_len = int(e["maxEpisodeLength"])
        self.len_obs = int(e.get("lenObservations", 1))
        self.len_states = int(e.get("lenStates", self.len_obs))
        self.cfi = int(e.get("controlFrequencyInv", 10))
        self.clip_obs = float(e.get("clipObservations", np.inf))
        self.clip_states = float(e.get("clipStates", np.inf))
        self.clip_actions = float(e.get("clipActions", np.inf))
        # gymapi.SimParams.dt is a C float: the python value is float32(dt) (FPV:219)
        self.dt = float(np.float32(cfg["sim"]["dt"]))
        self.substeps = int(cfg["sim"].get("substeps", 2))           # VT:432
        for k in ("random_copter_pos", "random_copter_quat", "random_copter_vel", "random_target_pos",
                  "random_target_yaw", "battery_consumption", "random_voltage", "rotor_response_time",
                  "rotor_noise", "rotor_delay", "rotor_response", "random_rotordynamic_coe",
                  "random_rotor_delay", "random_rotor_response", "random_rotor_speed",
                  "random_aerodynamic_coe", "delay_time_max", "delay_time", "ramdom_delay_time",
                  "ramdom_deploy_time", "random_command", "observation_noise", "difficulty"):
            setattr(self, k, cfg[k])                                   # FPV:63-112
        assert self.delay_time_max == 100, "arange(100) is hard-coded at FPV:329"
        N = self.N
        # per-env task id; mix = contiguous thirds of the GLOBAL env range (FPV:924-926)
        if self.task_mode == TASK_MIX:
            n1 = int(self.N_global / 3 * 1)
            n2 = int(self.N_global / 3 * 2)
            g = self.gid.astype(np.int64)
            self.task = torch.from_numpy(np.where(g < n1, TASK_POS, np.where(g < n2, TASK_ROTATE, TASK_FLIP)))
        else:
            self.task = torch.full((N,), self.task_mode, dtype=torch.int64)
        z = lambda *s: torch.zeros(*s, dtype=torch.float32)
        # root states: default pose (0,0,4), identity, at rest (FPV:264-266)
        self.pos = z(N, 3); self.pos[:, 2] = 4.0
        self.quat = z(N, 4); self.quat[:, 3] = 1.0
        self.linvel = z(N, 3); self.angvel = z(N, 3)
        self.tpos = z(N, 3); self.tpos[:, 2] = 4.0
        self.tquat = z(N, 4); self.tquat[:, 3] = 1.0
        self.rpy = torch.stack(euler_xyz(self.quat), dim=1)            # FPV:133-136
        self.rpy_old = self.rpy.clone()
        self.rpy_cont = self.rpy.clone()
        self.first_reset = True                                        # FPV:155-157
        self.command = z(N, 2)                                         # FPV:151
        self.flip_radian = z(N)                                        # FPV:828,927
        self.pid_prev = z(N, 3)                                        # angvel_control.py:64
        self.bat_u1, self.bat_ec, self.bat_t = z(N, 1), z(N, 1), z(N, 1)
        self.poly = f32(dyn.OMEGA_POLY).repeat(N, 1)
        self.lag_gain = dyn.ROTOR_SAMPLE_TIME / (self.rotor_response_time * torch.ones(N, 4))
        self.aero = f32((dyn.AERO_KF, dyn.AERO_KT, dyn.AERO_DX, dyn.AERO_DY, dyn.AERO_KTH)).repeat(N, 1)
        self.rotor_speed = z(N, 4)                                     # FPV:174
        self.volt = z(N, 1)
        self.actions, self.actions_old = z(N, 4), z(N, 4)              # FPV:185-186
        self.delay_buf = z(N, 4, 100)                                  # FPV:189
        if self.ramdom_delay_time:                                     # FPV:190-193 (drawn at step -1)
            self.delay_len = self._delay_draw(np.uint32(0xFFFFFFFF))
        else:
            self.delay_len = torch.full((N,), int(self.delay_time), dtype=torch.int64)
        # VT:240-253
        self.obs_buf = z(N, self.len_obs, 26)
        self.states_buf = z(N, self.len_states, 26)
        self.rew_buf = z(N)
        self.reset_buf = torch.ones(N, dtype=torch.int64)
        self.timeout_buf = torch.zeros(N, dtype=torch.bool)
        self.progress_buf = torch.zeros(N, dtype=torch.int64)
        self.step_index = 0
        self.ep_return = z(N)
        self.overflow = torch.zeros(N, dtype=torch.bool)
        self.stats = dict(sum_reward=0.0, n_done=0, n_timeout=0, sum_ep_return=0.0, sum_ep_len=0.0,
                          n_nonfinite=0, n_delay_overflow=0, n_steps=0)
        self.last_delay_index = None

    # ------------------------------------------------------------------ RNG helpers
    def _block(self, slot, stream):
        return px.draw(self.seed, self.gid, self.step_index & 0xFFFFFFFF, slot, stream)

    def _delay_draw(self, step):
        """FPV:191,576: clamp(delay - clamp(round(N(0,1)), -3, 3), min=0)."""
        w = px.draw(self.seed, self.gid, step, 1, px.STREAM_RESET)[:, 3]
        return torch.from_numpy(np.maximum(int(self.delay_time) - px.round_normal(w, 3), 0).astype(np.int64))

    # ------------------------------------------------------------------ resets
    def _reset(self, R):
        """FPV:475-517 for the envs in boolean mask R, plus the command re-draw for
        R | progress==500 (FPV:587-603)."""
        d = float(self.difficulty)
        N = self.N
        cmd_mask = R | (self.progress_buf == 500)                      # FPV:593-601 (before counters clear)
        if bool(R.any()):
            blk = [self._block(s, px.STREAM_RESET) for s in range(N_RESET_SLOTS)]
            U = lambda s, w: f32(px.u01(blk[s][:, w]))
            is_flip = self.task == TASK_FLIP
            mix = self.task_mode == TASK_MIX
            Rc = R.unsqueeze(1)
            # ---- copter (FPV:725-756 pos, :783-812 rotate, :850-884 flip, :981-1056 mix)
            if self.random_copter_pos:
                wide_xy = torch.stack((rand_range(-2, 2, U(0, 0)), rand_range(-2, 2, U(0, 1))), dim=1)
                wide_z = 2.5 + rand_range(-2, 2, U(0, 2))
                lim = 0.5 + 1.5 * d
                flip_xy = torch.stack((rand_range(-lim, lim, U(0, 0)), rand_range(-lim, lim, U(0, 1))), dim=1)
                flip_z = 3 + d * rand_range(-2, 2, U(0, 2))
                if mix:                                                # FPV:993-995,1012-1014,1031-1033
                    new_xy, new_z = wide_xy, wide_z
                else:
                    new_xy = torch.where(is_flip.unsqueeze(1), flip_xy, wide_xy)
                    new_z = torch.where(is_flip, flip_z, wide_z)
            else:
                small_xy = torch.stack((rand_range(-0.5, 0.5, U(0, 0)), rand_range(-0.5, 0.5, U(0, 1))), dim=1)
                if mix:                                                # FPV:997-998 etc.
                    new_xy, new_z = torch.zeros(N, 2), torch.full((N,), 2.5)
                else:
                    t = self.task_mode
                    new_xy = torch.zeros(N, 2) if t == TASK_POS else small_xy      # FPV:735-737 / :792 / :860
                    new_z = torch.full((N,), 3.0 if t == TASK_FLIP else 2.5)
            self.pos = torch.where(Rc, torch.cat((new_xy, new_z.unsqueeze(1)), dim=1), self.pos)
            if self.random_copter_quat:
                # rand_quat(n, pitch_lim, roll_lim, yaw_lim): the "pitch" draw lands in the ROLL slot (FPV:698-704)
                full = quat_from_euler(rand_range(-math.pi, math.pi, U(1, 0)), rand_range(-math.pi, math.pi, U(1, 1)),
                                       rand_range(-math.pi, math.pi, U(1, 2)), self._sincos)
                zero = torch.zeros(N)
                roll_only = quat_from_euler(rand_range(-math.pi, math.pi, U(1, 0)), zero, zero, self._sincos)      # FPV:864,1038
                new_q = torch.where(is_flip.unsqueeze(1), roll_only, full)
            else:
                new_q = torch.tensor([0.0, 0.0, 0.0, 1.0]).repeat(N, 1)
            self.quat = torch.where(Rc, new_q, self.quat)
            if self.random_copter_vel:
                lin_a = 3 * torch.stack([rand_range(-1.0, 1.0, U(2, i)) for i in range(3)], dim=1)
                ang_a = 3 * torch.stack([rand_range(-1.0, 1.0, U(3, i)) for i in range(3)], dim=1)
                lin_f = torch.stack([rand_range(-3 * d, 3 * d, U(2, i)) for i in range(3)], dim=1)
                coin = torch.from_numpy(np.where(blk[0][:, 3] >> np.uint32(31), 1.0, -1.0).astype(np.float32))
                ang_f = self.angvel.clone()                            # w_y, w_z stay stale (FPV:876,1047)
                ang_f[:, 0] = 10 * coin
                new_lin = torch.where(is_flip.unsqueeze(1), lin_f, lin_a)
                new_ang = torch.where(is_flip.unsqueeze(1), ang_f, ang_a)
            else:
                new_lin = torch.zeros(N, 3)
                # FpvFlip leaves angvel untouched (FPV:877-878); FpvMix zeroes it (FPV:1048-1050)
                keep = is_flip & (not mix)
                new_ang = torch.where(keep.unsqueeze(1), self.angvel, torch.zeros(N, 3))
            self.linvel = torch.where(Rc, new_lin, self.linvel)
            self.angvel = torch.where(Rc, new_ang, self.angvel)
            rpy_new = torch.stack(euler_xyz(self.quat), dim=1)         # FPV:752-754
            self.rpy = torch.where(Rc, rpy_new, self.rpy)
            self.rpy_old = torch.where(Rc, rpy_new, self.rpy_old)
            self.rpy_cont = torch.where(Rc, rpy_new, self.rpy_cont)
            # ---- controllers (FPV:550-558)
            self.pid_prev = torch.where(Rc, torch.zeros(N, 3), self.pid_prev)      # angvel_control.py:90-94
            ec0 = rand_range(0, 2.2, U(2, 3)).unsqueeze(1) if self.random_voltage else torch.zeros(N, 1)
            self.bat_u1 = torch.where(Rc, torch.zeros(N, 1), self.bat_u1)          # battery_dynamics.py:38-45
            self.bat_ec = torch.where(Rc, ec0, self.bat_ec)
            self.bat_t = torch.where(Rc, torch.zeros(N, 1), self.bat_t)
            lo, hi = 1 - 0.05 * d, 1 + 0.05 * d
            nominal_poly = f32(dyn.OMEGA_POLY).repeat(N, 1)
            if self.random_rotordynamic_coe:                           # thrust_dynamics.py:117-122
                fac = torch.stack([rand_range(lo, hi, U(5, 0)), rand_range(lo, hi, U(5, 1)), rand_range(lo, hi, U(5, 2)),
                                   rand_range(lo, hi, U(5, 3)), rand_range(lo, hi, U(6, 0))], dim=1)
                new_poly = nominal_poly * fac
            else:
                new_poly = nominal_poly
            self.poly = torch.where(Rc, new_poly, self.poly)
            tau0 = float(self.rotor_response_time)
            if self.rotor_response:                                    # thrust_dynamics.py:134-141
                if self.random_rotor_response:
                    tau = torch.stack([rand_range(tau0 - 0.001, tau0 + 0.001, U(8, i)) for i in range(4)], dim=1)
                else:
                    tau = tau0 * torch.ones(N, 4)
            else:
                tau = dyn.ROTOR_SAMPLE_TIME * torch.ones(N, 4)
            self.lag_gain = torch.where(Rc, dyn.ROTOR_SAMPLE_TIME / tau, self.lag_gain)
            if self.random_rotor_speed:                                # thrust_dynamics.py:143-146
                w0 = torch.stack([rand_range(0, 400, U(9, i)) for i in range(4)], dim=1)
            else:
                w0 = torch.zeros(N, 4)
            self.rotor_speed = torch.where(Rc, w0, self.rotor_speed)
            if self.random_aerodynamic_coe:                            # thrust_dynamics.py:201-210
                nominal_aero = f32((dyn.AERO_KF, dyn.AERO_KT, dyn.AERO_DX, dyn.AERO_DY, dyn.AERO_KTH)).repeat(N, 1)
                fac = torch.stack([rand_range(lo, hi, U(6, 1)), rand_range(lo, hi, U(6, 2)), rand_range(lo, hi, U(6, 3)),
                                   rand_range(lo, hi, U(7, 0)), rand_range(lo, hi, U(7, 1))], dim=1)
                self.aero = torch.where(Rc, nominal_aero * fac, self.aero)
            # ---- env signals (FPV:560-581)
            self.volt = torch.where(Rc, torch.zeros(N, 1), self.volt)
            self.actions = torch.where(Rc, torch.zeros(N, 4), self.actions)
            self.actions_old = torch.where(Rc, torch.zeros(N, 4), self.actions_old)
            self.delay_buf = torch.where(R.view(N, 1, 1), torch.zeros(N, 4, 100), self.delay_buf)
            if self.ramdom_delay_time:
                new_len = self._delay_draw(self.step_index & 0xFFFFFFFF)
            else:
                new_len = torch.full((N,), int(self.delay_time), dtype=torch.int64)
            self.delay_len = torch.where(R, new_len, self.delay_len)
            self.overflow = self.overflow & ~R
            # ---- target (FPV:523-548)
            if self.random_target_pos:
                txy = d * torch.stack((rand_range(-2, 2, U(4, 0)), rand_range(-2, 2, U(4, 1))), dim=1)
                tz = 3 + d * rand_range(-2, 2, U(4, 2))
            else:
                txy, tz = torch.zeros(N, 2), torch.full((N,), 3.0)
            self.tpos = torch.where(Rc, torch.cat((txy, tz.unsqueeze(1)), dim=1), self.tpos)
            yaw = rand_range(-math.pi, math.pi, U(3, 3)) if self.random_target_yaw else torch.zeros(N)
            zero = torch.zeros(N)
            self.tquat = torch.where(Rc, quat_from_euler(zero, zero, yaw, self._sincos), self.tquat)
        # ---- command (FPV:758-759, :814-821, :886-917, :1058-1112)
        if bool(cmd_mask.any()):
            cblk = self._block(0, px.STREAM_COMMAND)
            at500 = self.progress_buf == 500
            is_pos, is_rot, is_flip = (self.task == TASK_POS), (self.task == TASK_ROTATE), (self.task == TASK_FLIP)
            m = cmd_mask & is_pos
            self.command = torch.where(m.unsqueeze(1), torch.zeros(N, 2), self.command)
            m = cmd_mask & is_rot
            speed = rand_range(-6, 6, f32(px.u01(cblk[:, 1]))) if self.random_command else torch.ones(N)
            self.command[:, 0] = torch.where(m, torch.ones(N), self.command[:, 0])
            self.command[:, 1] = torch.where(m, speed, self.command[:, 1])
            # flip: eighths of U(0,1) -> {-3,-2,-1,0,0,1,2,3} turns (FPV:892-901)
            turns = torch.from_numpy(np.array([-3, -2, -1, 0, 0, 1, 2, 3], dtype=np.float32)[(cblk[:, 0] >> np.uint32(29)).astype(np.int64)])
            m = at500 & is_flip
            self.flip_radian = torch.where(m, self.flip_radian + 2 * math.pi * turns, self.flip_radian)
            m = R & is_flip
            first = torch.where(self.angvel[:, 0] > 5, torch.full((N,), TWO_PI), torch.full((N,), -TWO_PI))   # FPV:913
            self.flip_radian = torch.where(m, first, self.flip_radian)
            if self.task_mode == TASK_FLIP:
                self.command[:, 0] = -1                                # FPV:917 (every env)
            else:
                m = cmd_mask & is_flip
                self.command[:, 0] = torch.where(m, -torch.ones(N), self.command[:, 0])              # FPV:1112
        self.reset_buf = torch.where(R, torch.zeros_like(self.reset_buf), self.reset_buf)            # FPV:510-511
        self.progress_buf = torch.where(R, torch.zeros_like(self.progress_buf), self.progress_buf)

    # ------------------------------------------------------------------ state refresh
    def _refresh(self, w_body=None):
        """FPV:334-360.  ``w_body``: body-frame angular velocity carried by the simulator stand-in between the
        control sub-steps of one RL step (oracle/rigid_body.py); None = derive it from the world-frame root
        state exactly as FPV:350 does (first sub-step and post_physics_step)."""
        self.rpy = torch.stack(euler_xyz(self.quat), dim=1)
        delta = self.rpy - self.rpy_old
        if self.first_reset:
            self.first_reset = False
        else:
            delta = torch.where(delta > 1, delta - TWO_PI, delta)
            delta = torch.where(delta < -1, delta + TWO_PI, delta)
        self.rpy_cont = self.rpy_cont + delta
        self.rpy_old = self.rpy.clone()
        qc = qconj(self.quat)
        self.linvel_body = qrot(qc, self.linvel)
        self.angvel_body = qrot(qc, self.angvel) if w_body is None else w_body
        self.rel_pos = self.tpos - self.pos
        self.rel_pos_body = qrot(qc, self.rel_pos)
        self.rel_quat_body = qmul(qc, self.tquat)
        self.rel_linvel = torch.zeros_like(self.linvel) - self.linvel   # target velocities are never written (FPV:147-148)
        self.rel_angvel = torch.zeros_like(self.angvel) - self.angvel
        self.rel_linvel_body = qrot(qc, self.rel_linvel)
        self.rel_angvel_body = qrot(qc, self.rel_angvel)

    # ------------------------------------------------------------------ the step
    def step(self, actions):
        """VT:290-334.  actions: (N,4) float32 torch tensor."""
        N = self.N
        a = torch.clamp(actions.to(torch.float32), -self.clip_actions, self.clip_actions)          # VT:304
        # ---- pre_physics_step, FPV:317-332
        R = self.reset_buf != 0
        self._reset(R)
        self.actions_old = self.actions.clone()
        self.actions = a.clone()
        if self.ramdom_deploy_time:                                    # FPV:323-324
            w = self._block(0, px.STREAM_DEPLOY)[:, 0]
            T = torch.from_numpy((10 - px.round_normal(w, 1)).astype(np.int64))
        else:
            T = torch.full((N,), 10, dtype=torch.int64)
        slots = torch.arange(100).view(1, 1, 100)
        start = self.delay_len.view(N, 1, 1)
        mask = (slots >= start) & (slots < start + T.view(N, 1, 1))
        self.delay_buf = torch.where(mask, a.unsqueeze(-1).expand(N, 4, 100), self.delay_buf)       # FPV:327-330
        self.overflow = self.overflow | ((self.delay_len + T) > 100)
        self.delay_len = self.delay_len + T
        rows = torch.arange(N)
        delay_idx_log, delay_act_log = [], []
        # ---- control_freq_inv x (mid_physics_step + simulate), VT:309-313
        w_b = None
        for k in range(self.cfi):
            self._refresh(w_b)                                         # FPV:363
            idx = torch.clamp(self.delay_len - 1, max=k)               # FPV:366 (negative wraps like python)
            idx = torch.where(idx < 0, idx + 100, idx)
            delay_idx_log.append(idx.clone())
            da = self.delay_buf[rows, :, idx]
            delay_act_log.append(da.clone())
            # angular_vel_control, FPV:637-650
            u0 = (da[:, 0] + 1) / 2 * 1000
            sp = da[:, 1:] * 20
            trq, self.pid_prev = dyn.rate_pid(sp, self.angvel_body, self.pid_prev, self.dt)
            u = torch.cat((u0.unsqueeze(1), trq), dim=1)
            throttle = dyn.allocate(u)
            # control_with_thrusts, FPV:608-635
            p_m = dyn.mech_power(self.rotor_speed)
            self.volt, self.bat_u1, self.bat_ec, self.bat_t = dyn.battery_step(
                p_m, self.bat_u1, self.bat_ec, self.bat_t, self.dt, bool(self.battery_consumption))
            self.rotor_speed = dyn.rotor_step(self.volt, throttle, self.rotor_speed, self.poly, self.lag_gain)
            if self.rotor_noise:                                       # thrust_dynamics.py:68-78
                ratio = 10 / 700
                nb = self._block(k, px.STREAM_ROTOR_NOISE)
                self.rotor_speed = self.rotor_speed * rand_range(1 - ratio, 1 + ratio, f32(px.u01(nb)))
            f, tq, body_f = dyn.aero_step(self.linvel_body, self.rotor_speed, self.aero)
            fs, ts = dyn.real_to_sim(f, tq)
            force_b, torque_b = rb.body_wrench(fs, ts, body_f)
            force_b = torch.where(R.unsqueeze(1), torch.zeros(N, 3), force_b)                      # FPV:629-630
            torque_b = torch.where(R.unsqueeze(1), torch.zeros(N, 3), torque_b)
            self.pos, self.quat, self.linvel, w_b = rb.integrate(                                  # VT:313
                self.pos, self.quat, self.linvel, self.angvel_body, force_b, torque_b, self.dt, self.substeps)
            if self.reference_exact:                                   # the simulator hands back a world-frame root state
                self.angvel = qrot(self.quat, w_b)
                w_b = None
        if w_b is not None and not self.reference_exact:
            self.angvel = qrot(self.quat, w_b)                         # world-frame root state at the end of the RL step
        self.last_delay_index = torch.stack(delay_idx_log, dim=1)
        self.last_delayed_actions = torch.stack(delay_act_log, dim=1)          # (N, cfi, 4)
        # ---- post_physics_step, FPV:374-388
        self.progress_buf = self.progress_buf + 1
        self.delay_buf[:, :, 0:-10] = self.delay_buf[:, :, 10:].clone()                             # memmove semantics
        self.delay_len = torch.clamp(self.delay_len - 10, min=0)
        self._refresh()
        self._observe()
        self._reward()
        self.timeout_buf = (self.progress_buf >= self.max_len - 1) & (self.reset_buf != 0)         # VT:323
        self._episode_stats()
        self.step_index += 1
        obs = {"obs": torch.clamp(self.obs_buf, -self.clip_obs, self.clip_obs),
               "states": torch.clamp(self.states_buf, -self.clip_states, self.clip_states)}        # VT:331-332
        return obs, self.rew_buf, self.reset_buf, {"time_outs": self.timeout_buf}

    def reset(self):
        """VT:352-361: returns the (zero) buffers, no simulation."""
        return {"obs": torch.clamp(self.obs_buf, -self.clip_obs, self.clip_obs),
                "states": torch.clamp(self.states_buf, -self.clip_states, self.clip_states)}

    # ------------------------------------------------------------------ observation
    def _frame(self, rel_quat_body):
        """Newest 26-value frame, FPV:394-400 / :415-421 + task override."""
        N = self.N
        fr = torch.zeros(N, 26)
        fr[:, 0:3] = self.rel_pos_body / 3
        fr[:, 3:12] = rotmat9(rel_quat_body)
        fr[:, 12:15] = self.rel_linvel_body / 2
        fr[:, 15:18] = self.rel_angvel_body / math.pi
        fr[:, 18] = (self.volt.squeeze(1) - 23) / 3
        fr[:, 19:23] = self.actions
        fr[:, 23] = 4 * torch.clamp(self.pos[:, 2], 0, 0.5) - 1
        return fr

    def _observe(self):
        d = float(self.difficulty)
        is_rot, is_flip = (self.task == TASK_ROTATE), (self.task == TASK_FLIP)
        # flip: command1 <- clamp(flip_radian - roll_continuous, +-2pi)  (FPV:831-832, :930-931)
        c1 = torch.clamp(self.flip_radian - self.rpy_cont[:, 0], min=-TWO_PI, max=TWO_PI)
        self.command[:, 1] = torch.where(is_flip, c1, self.command[:, 1])
        clean = self._frame(self.rel_quat_body)
        noisy = clean.clone()
        if self.observation_noise:                                     # FPV:402-410
            nrm = []
            for s in range(3):
                b = self._block(s, px.STREAM_OBS_NOISE)
                z0, z1 = px.box_muller(b[:, 0], b[:, 1])
                z2, z3 = px.box_muller(b[:, 2], b[:, 3])
                nrm += [f32(z0), f32(z1), f32(z2), f32(z3)]
            ub = self._block(3, px.STREAM_OBS_NOISE)
            sig_p, sig_v, sig_w, sig_u, sig_h = 0.06 / 3 / 3, 0.1 / 3 / 2, 60 / 3 / 180, 0.06 / 3, 0.06 / 3 / 3
            for i in range(3):
                noisy[:, i] = noisy[:, i] + d * (nrm[i] * sig_p)
            lim = d * 0.05
            nq = quat_from_euler(rand_range(-lim, lim, f32(px.u01(ub[:, 0]))), rand_range(-lim, lim, f32(px.u01(ub[:, 1]))),
                                 rand_range(-lim, lim, f32(px.u01(ub[:, 2]))), self._sincos)
            noisy[:, 3:12] = rotmat9(qmul(self.rel_quat_body, nq))
            for i in range(3):
                noisy[:, 12 + i] = noisy[:, 12 + i] + d * (nrm[3 + i] * sig_v)
                noisy[:, 15 + i] = noisy[:, 15 + i] + d * (nrm[6 + i] * sig_w)
            noisy[:, 18] = noisy[:, 18] + d * (nrm[9] * sig_u)
            noisy[:, 23] = noisy[:, 23] + d * (nrm[10] * sig_h)
        # task id / command (FPV:713-714, :768-771, :835-838)
        scale = torch.where(is_rot, torch.full_like(c1, 6.0), torch.ones_like(c1))
        cmd1 = torch.where(is_flip, self.command[:, 1] / 2 / math.pi, self.command[:, 1] / scale)
        for fr in (clean, noisy):
            fr[:, 24] = self.command[:, 0]
            fr[:, 25] = cmd1
        self.obs_buf[:, :-1, :] = self.obs_buf[:, 1:, :].clone()       # FPV:392 (history is never cleared on reset)
        self.obs_buf[:, -1, :] = noisy
        self.states_buf[:, :-1, :] = self.states_buf[:, 1:, :].clone()  # FPV:413
        self.states_buf[:, -1, :] = clean
        self.stats["n_nonfinite"] += int((~torch.isfinite(clean).all(dim=1)).sum())

    # ------------------------------------------------------------------ reward
    def _reward(self):
        """FPV:716-723, :773-781, :841-848, :948-979.  Task groups are contiguous index
        ranges (single task: one range; mix: thirds, FPV:924-926), evaluated per range
        exactly like the reference's slicing."""
        rew = torch.zeros(self.N)
        reset = torch.zeros(self.N, dtype=torch.int64)
        for task, lo, hi in self._task_ranges():
            if hi <= lo:
                continue
            sl = slice(lo, hi)
            if task == TASK_POS:
                r, x = rw.pos_reward(self.rel_pos_body[sl], self.pos[sl], self.quat[sl], self.tquat[sl],
                                     self.progress_buf[sl], self.max_len)
            elif task == TASK_ROTATE:
                r, x = rw.rotate_reward(self.rel_pos[sl], self.rel_linvel[sl], self.pos[sl], self.quat[sl],
                                        self.command[sl], self.progress_buf[sl], self.max_len)
            else:
                r, x = rw.flip_reward(self.rel_pos_body[sl], self.rel_quat_body[sl], self.pos[sl], self.command[sl],
                                      self.progress_buf[sl], self.max_len)
            rew[sl] = r
            reset[sl] = x
        self.rew_buf, self.reset_buf = rew, reset

    def _task_ranges(self):
        if self.task_mode != TASK_MIX:
            return [(self.task_mode, 0, self.N)]
        t = self.task
        n_pos = int((t == TASK_POS).sum())
        n_rot = int((t == TASK_ROTATE).sum())
        return [(TASK_POS, 0, n_pos), (TASK_ROTATE, n_pos, n_pos + n_rot), (TASK_FLIP, n_pos + n_rot, self.N)]

    def _episode_stats(self):
        """Device-side equivalent of the rollout bookkeeping at ppo_asymmetry.py:313-339."""
        done = self.reset_buf != 0
        self.ep_return = self.ep_return + self.rew_buf
        s = self.stats
        s["sum_reward"] += float(self.rew_buf.double().sum())
        s["n_done"] += int(done.sum())
        s["n_timeout"] += int((self.timeout_buf & done).sum())
        s["sum_ep_return"] += float(self.ep_return[done].double().sum())
        s["sum_ep_len"] += float(self.progress_buf[done].double().sum())
        s["n_delay_overflow"] += int(self.overflow.sum())
        s["n_steps"] += self.N
        self.ep_return = torch.where(done, torch.zeros_like(self.ep_return), self.ep_return)

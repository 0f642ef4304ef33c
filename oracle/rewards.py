"""Task rewards + termination of the TACO fpv tasks, torch-float32 CPU restatement.
TEST INFRASTRUCTURE.  Reference: IsaacGymEnvs/isaacgymenvs/tasks/control/task_reward.py
(RW).  Golden vectors from the reference TorchScript functions: tests/golden/rewards.npz.
"""
import math

import torch

from .leaf_math import quat_angle, rotmat9, cross3


def _two_scale(d):
    """1/(1+d^2) + 1/(1+10 d^2), the recurring kernel of RW:26-28,33-35,69-71 ...
    (note the reference's evaluation order: 10 * d * d == (10*d)*d)."""
    return 1.0 / (1.0 + d * d) + 1.0 / (1.0 + 10 * d * d)


def _termination(z, dist, progress, max_len):
    """RW:39-45 (identical in all three rewards): die if z < 0.1 or dist > 10; any env at
    progress >= max_len-1 resets.  int64 like reset_buf (vec_task_asymmetry.py:246-247)."""
    die = torch.zeros_like(progress)
    one = torch.ones_like(progress)
    die = torch.where(z < 0.1, one, die)
    die = torch.where(dist > 10, one, die)
    return torch.where(progress >= max_len - 1, one, die)


def pos_reward(rel_pos_body, copter_pos, copter_quat, target_quat, progress, max_len):
    """RW:20-47 (compute_pos_reward)."""
    dist = torch.norm(rel_pos_body, dim=1)
    ang = quat_angle(copter_quat, target_quat)
    rew = _two_scale(dist) * _two_scale(ang)
    return rew / 100, _termination(copter_pos[:, 2], dist, progress, max_len)


def rotate_reward(rel_pos, rel_linvel, copter_pos, copter_quat, command, progress, max_len):
    """RW:50-104 (compute_rotating_reward): circle of radius 1.2 m around the target."""
    radius = 1.2
    speed = command[:, -1]
    e_z = torch.zeros_like(rel_pos)
    e_z[:, 2] = 1
    e_x = -rel_pos
    e_x[:, 2] = 0
    e_x = e_x / (torch.norm(e_x, dim=1, keepdim=True) + 1e-8)
    e_y = cross3(e_z, e_x)
    e_y = e_y / (torch.norm(e_y, dim=1, keepdim=True) + 1e-8)

    hori = torch.norm(rel_pos[:, :2], dim=1) - radius
    vert = torch.abs(rel_pos[:, 2])
    dist = torch.sqrt(hori ** 2 + vert ** 2)
    r_pos = _two_scale(dist)

    want = torch.zeros_like(rel_linvel)
    want[:, 1] = speed
    v_n = torch.sum(rel_linvel * e_x, dim=1, keepdim=True)
    v_t = torch.sum(rel_linvel * e_y, dim=1, keepdim=True)
    v_new = torch.cat((v_n, v_t, rel_linvel[:, 2:3]), dim=1)
    v_err = torch.norm(v_new - want, dim=1)
    r_vel = _two_scale(v_err)

    heading = rotmat9(copter_quat).reshape(-1, 3, 3)[:, :, 0]
    d_dir = 1 + torch.sum(e_x[:, :2] * heading[:, :2], dim=1) / torch.norm(heading[:, :2], dim=1)
    r_dir = _two_scale(d_dir)

    rew = r_pos * r_vel * r_dir
    return rew / 100, _termination(copter_pos[:, 2], dist, progress, max_len)


def flip_reward(rel_pos_body, rel_quat_body, copter_pos, command, progress, max_len):
    """RW:107-143 (compute_flip_reward)."""
    dist = torch.norm(rel_pos_body, dim=1)
    r_pos = 1.0 / (1.0 + 1 * dist) + 1.0 / (1.0 + 10 * dist)
    tilt = 1 - rotmat9(rel_quat_body)[:, 0]
    r_tilt = 1.0 / (1.0 + 10 * tilt)
    turns = command[:, -1] / 2 / math.pi
    r_cmd = _two_scale(turns)
    rew = r_pos * r_tilt * r_cmd
    return rew / 100, _termination(copter_pos[:, 2], dist, progress, max_len)

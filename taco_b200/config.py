"""Reconstructed default task configuration for the fpv_asymmetry tasks.

The reference loads ``isaacgymenvs/cfg/Fpv_asymmetry_PPO_<task>.yaml``
(IsaacGymEnvs/train/train_fpv_asymmetry_ppo.py:257-271) but the YAML files were
git-ignored upstream and are absent.  ``make_cfg`` rebuilds the ``cfg['Task']`` dict the
env constructor receives, with every key the reference reads
(IsaacGymEnvs/isaacgymenvs/tasks/fpv_asymmetry.py:57-115,
tasks/base/vec_task_asymmetry.py:64-100,423-456), including the upstream spellings
``ramdom_delay_time`` / ``ramdom_deploy_time``.  Values the code pins: dt = 0.001,
controlFrequencyInv = 10, delay_time_max = 100; the rest are the
README training settings (README.md:60-66) or marked as our choice.
"""
import copy

TASK_MODES = ("pos", "rotate", "flip", "mix")

_DEFAULT = {
    "name": "Fpv",
    "task_mode": "flip",
    "seed": 0,
    "physics_engine": "physx",
    "env": {
        "numEnvs": 4096,                # README.md:41
        "envSpacing": 5.0,              # unused by the B200 path (no viewer)
        "maxEpisodeLength": 1000,       # train_fpv_asymmetry_ppo.py:342
        "enableDebugVis": False,
        "lenObservations": 1,           # README.md:60
        "lenStates": 5,                 # README.md:60
        "controlFrequencyInv": 10,      # pinned by fpv_asymmetry.py:326,378
        "clipObservations": float("inf"),
        "clipStates": float("inf"),
        "clipActions": float("inf"),
    },
    "sim": {
        "dt": 0.001,                    # pinned by thrust_dynamics.py:34
        "substeps": 2,                  # vec_task_asymmetry.py:432 default
        "up_axis": "z",
        "gravity": [0.0, 0.0, -9.81],
        "use_gpu_pipeline": True,
        "physx": {},
    },
    "task": {"randomization_params": {}},
    "random_copter_pos": True,
    "random_copter_quat": True,
    "random_copter_vel": True,
    "random_target_pos": True,
    "random_target_yaw": True,
    "battery_consumption": True,
    "random_voltage": True,
    "rotor_response_time": 0.017,       # README.md:60
    "rotor_noise": False,
    "rotor_delay": True,
    "rotor_response": True,
    "random_rotordynamic_coe": False,
    "random_rotor_delay": False,
    "random_rotor_response": False,
    "random_rotor_speed": True,
    "random_aerodynamic_coe": False,
    "delay_time_max": 100,              # forced by fpv_asymmetry.py:329
    "delay_time": 20,                   # README.md:60
    "ramdom_delay_time": False,
    "ramdom_deploy_time": False,
    "random_command": True,
    "observation_noise": False,
    "difficulty": 1.0,
    "record_flag": False,
    "record_path": "",
}


def make_cfg(task_mode="flip", num_envs=4096, domain_randomization=False, **overrides):
    """Build a reference-style cfg dict.  ``domain_randomization=True`` turns on every
    per-env randomisation switch (BASELINE.json config 5).  Keyword overrides may name
    top-level keys or ``env.<key>`` / ``sim.<key>``."""
    if task_mode not in TASK_MODES:
        raise ValueError(f"unknown task_mode {task_mode!r}; expected one of {TASK_MODES}")
    cfg = copy.deepcopy(_DEFAULT)
    cfg["task_mode"] = task_mode
    cfg["name"] = "Fpv_" + task_mode    # train_fpv_asymmetry_ppo.py:287
    cfg["env"]["numEnvs"] = int(num_envs)
    if domain_randomization:
        for k in ("random_rotor_response", "ramdom_delay_time", "ramdom_deploy_time", "random_voltage",
                  "random_rotordynamic_coe", "random_aerodynamic_coe"):
            cfg[k] = True
    for k, v in overrides.items():
        if "." in k:
            sec, key = k.split(".", 1)
            cfg[sec][key] = v
        else:
            cfg[k] = v
    return cfg

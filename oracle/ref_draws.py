"""Feed the REFERENCE'S OWN env classes (run over oracle/fake_gym.py) the same random numbers the oracle and the
CUDA kernel draw, so that their trajectories can be compared draw for draw.  TEST INFRASTRUCTURE, build
container only.

The reference takes every random number from torch's global generator (``torch_rand_float`` -> ``torch.rand``,
python/isaacgym/torch_utils.py:216-219; ``torch.rand`` / ``torch.normal`` in fpv_asymmetry.py), in call order over
index lists -- irreproducible by any other implementation.  Oracle and kernel use a counter-based Philox stream
keyed by (global env id, RL step, slot, stream) (oracle/philox.py, slot table in oracle/fpv_env.py).  This module
replaces, in the *module globals* of the loaded reference modules only (never the files, never torch itself),

    torch_rand_float      in fpv_asymmetry / control.thrust_dynamics / control.battery_dynamics
    torch (a proxy whose .rand / .normal are ours, everything else torch's) in fpv_asymmetry

with functions that identify the CALL SITE (reference file + line number of the calling frame) and return the
Philox words the slot table assigns to that site, for the env ids the calling frame is working on.  The table
below is therefore also the line-by-line statement of how the slot table maps onto the reference's draws.
Every value is transformed exactly as the reference would transform its own sample: uniform draws go through
``(upper - lower) * u + lower`` (torch_utils.py:219); ``clamp(round(normal))`` sites receive an integer-valued
sample; gaussian noise sites receive ``z * std``.
"""
import sys

import numpy as np
import torch

from . import philox as px

R, C, D, O, N_ = px.STREAM_RESET, px.STREAM_COMMAND, px.STREAM_DEPLOY, px.STREAM_OBS_NOISE, px.STREAM_ROTOR_NOISE
UNUSED = "unused"          # a draw whose value the reference never reads (throttle_para, fpv_asymmetry.py:909,1105)

# (file tag, line) -> (stream, [(slot, word) per column], name of the env-id variable in the calling frame | None = all)
SITES = {
    # ---- fpv_asymmetry.py: FpvPos.reset_copter_idx
    ("fpv", 731): (R, [(0, 0)], "env_ids"), ("fpv", 732): (R, [(0, 1)], "env_ids"), ("fpv", 733): (R, [(0, 2)], "env_ids"),
    ("fpv", 746): (R, [(2, 0), (2, 1), (2, 2)], "env_ids"), ("fpv", 747): (R, [(3, 0), (3, 1), (3, 2)], "env_ids"),
    # FpvRotate.reset_copter_idx
    ("fpv", 789): (R, [(0, 0), (0, 1)], "env_ids"), ("fpv", 790): (R, [(0, 2)], "env_ids"),
    ("fpv", 792): (R, [(0, 0), (0, 1)], "env_ids"),
    ("fpv", 802): (R, [(2, 0), (2, 1), (2, 2)], "env_ids"), ("fpv", 803): (R, [(3, 0), (3, 1), (3, 2)], "env_ids"),
    # FpvFlip.reset_copter_idx
    ("fpv", 856): (R, [(0, 0), (0, 1)], "env_ids"), ("fpv", 857): (R, [(0, 2)], "env_ids"),
    ("fpv", 860): (R, [(0, 0), (0, 1)], "env_ids"),
    ("fpv", 870): (R, [(2, 0), (2, 1), (2, 2)], "env_ids"),
    ("fpv", 873): (R, [(0, 3)], None),                         # spin-direction coin, torch.rand(num_envs)
    # FpvMix.reset_copter_idx (pos / rotate / flip sub-lists)
    ("fpv", 994): (R, [(0, 0), (0, 1)], "env_ids0"), ("fpv", 995): (R, [(0, 2)], "env_ids0"),
    ("fpv", 1005): (R, [(2, 0), (2, 1), (2, 2)], "env_ids0"), ("fpv", 1006): (R, [(3, 0), (3, 1), (3, 2)], "env_ids0"),
    ("fpv", 1013): (R, [(0, 0), (0, 1)], "env_ids1"), ("fpv", 1014): (R, [(0, 2)], "env_ids1"),
    ("fpv", 1024): (R, [(2, 0), (2, 1), (2, 2)], "env_ids1"), ("fpv", 1025): (R, [(3, 0), (3, 1), (3, 2)], "env_ids1"),
    ("fpv", 1032): (R, [(0, 0), (0, 1)], "env_ids2"), ("fpv", 1033): (R, [(0, 2)], "env_ids2"),
    ("fpv", 1043): (R, [(2, 0), (2, 1), (2, 2)], "env_ids2"),
    ("fpv", 1044): (R, [(0, 3)], None),
    # reset_target_idx
    ("fpv", 531): (R, [(4, 0), (4, 1)], "env_ids"), ("fpv", 532): (R, [(4, 2)], "env_ids"), ("fpv", 543): (R, [(3, 3)], "env_ids"),
    # reset_command_idx: rotate speed, flip turn count at progress == 500
    ("fpv", 818): (C, [(0, 1)], "env_ids"), ("fpv", 1075): (C, [(0, 1)], "env_ids1"),
    ("fpv", 892): (C, [(0, 0)], None), ("fpv", 1088): (C, [(0, 0)], None),
    ("fpv", 909): UNUSED, ("fpv", 1105): UNUSED,
    # ---- control/battery_dynamics.py reset: E_c
    ("battery", 45): (R, [(2, 3)], "reset_env_ids"),
    # ---- control/thrust_dynamics.py: RotorDynamics.omega_noise / reset, AeroDynamics.reset
    ("thrust", 119): (R, [(5, 0), (5, 1), (5, 2), (5, 3), (6, 0)], "reset_env_ids"),
    ("thrust", 120): UNUSED,
    ("thrust", 137): (R, [(8, 0), (8, 1), (8, 2), (8, 3)], "reset_env_ids"),
    ("thrust", 144): (R, [(9, 0), (9, 1), (9, 2), (9, 3)], "reset_env_ids"),
    ("thrust", 207): (R, [(6, 1), (6, 2)], "reset_env_ids"), ("thrust", 209): (R, [(6, 3), (7, 0)], "reset_env_ids"),
    ("thrust", 210): (R, [(7, 1)], "reset_env_ids"),
}
# rand_quat (fpv_asymmetry.py:699-701 = word 0,1,2) by the line that CALLS it -> (stream, slot, env-id variable there)
RAND_QUAT_CALLERS = {740: (R, 1, "env_ids"), 796: (R, 1, "env_ids"), 864: (R, 1, "env_ids"),
                     1000: (R, 1, "env_ids0"), 1019: (R, 1, "env_ids1"), 1038: (R, 1, "env_ids2"),
                     405: (O, 3, None)}
# clamp(round(N(0,1))) sites -> (stream, slot, word, clip, step override, env-id variable)
ROUND_NORMAL_SITES = {191: (R, 1, 3, 3, 0xFFFFFFFF, None), 576: (R, 1, 3, 3, None, "env_ids"), 324: (D, 0, 0, 1, None, None)}
# gaussian observation noise sites -> index of the first normal in the 12-normal table of STREAM_OBS_NOISE
OBS_NORMAL_SITES = {403: 0, 407: 3, 408: 6, 409: 9, 410: 10}

_FILE_TAGS = {"fpv_asymmetry.py": "fpv", "thrust_dynamics.py": "thrust", "battery_dynamics.py": "battery"}


def _site(depth=2):
    """(tag, line, frame) of the nearest calling frame that lives in a reference file."""
    f = sys._getframe(depth)
    while f is not None:
        tag = _FILE_TAGS.get(f.f_code.co_filename.rsplit("/", 1)[-1])
        if tag:
            return tag, f.f_lineno, f
        f = f.f_back
    raise RuntimeError("random draw outside the reference files")


class Feeder:
    """One per reference env instance; ``activate()`` installs it into the reference modules' globals."""

    def __init__(self, seed, num_envs, env_offset=0):
        self.seed = int(seed)
        self.gid = np.arange(env_offset, env_offset + num_envs, dtype=np.uint64)
        self.N = num_envs
        self.step_index = 0
        self.env = None                               # set by the pinned subclass (needs mid_step_count)
        self.calls = 0

    # ---------------------------------------------------------------- helpers
    def _words(self, stream, slot, step=None):
        st = self.step_index & 0xFFFFFFFF if step is None else step
        return px.draw(self.seed, self.gid, st, slot, stream)

    def _ids(self, frame, var):
        if var is None:
            return np.arange(self.N)
        ids = frame.f_locals[var]
        ids = ids.numpy() if torch.is_tensor(ids) else np.asarray(ids, dtype=np.int64)
        return ids.astype(np.int64).reshape(-1)

    def _uniform(self, stream, cols, ids):
        out = np.empty((len(ids), len(cols)), dtype=np.float32)
        for j, (slot, word) in enumerate(cols):
            out[:, j] = px.u01(self._words(stream, slot)[ids, word])
        return torch.from_numpy(out)

    # ---------------------------------------------------------------- replacements
    def rand_float(self, lower, upper, shape, device):
        """Stands in for torch_utils.torch_rand_float (TU:216-219)."""
        self.calls += 1
        tag, line, frame = _site()
        shape = tuple(shape)
        if tag == "fpv" and line in (699, 700, 701):                       # rand_quat
            caller = frame.f_back
            stream, slot, var = RAND_QUAT_CALLERS[caller.f_lineno]
            u = self._uniform(stream, [(slot, line - 699)], self._ids(caller, var))
        elif tag == "thrust" and line == 75:                               # omega_noise, one slot per control sub-step
            k = int(self.env.mid_step_count) - 1
            u = self._uniform(N_, [(k, 0), (k, 1), (k, 2), (k, 3)], np.arange(self.N))
        else:
            spec = SITES[(tag, line)]
            if spec == UNUSED:
                u = torch.full(shape, 0.5)
            else:
                stream, cols, var = spec
                u = self._uniform(stream, cols, self._ids(frame, var))
        assert tuple(u.shape) == shape, (tag, line, tuple(u.shape), shape)
        return (upper - lower) * u + lower                                  # TU:219

    def rand(self, *size, **kw):
        """Stands in for torch.rand(num_envs) (fpv_asymmetry.py:873,892,909,1044,1088,1105)."""
        self.calls += 1
        tag, line, frame = _site()
        spec = SITES[(tag, line)]
        if spec == UNUSED:
            return torch.full(size, 0.5)
        stream, cols, var = spec
        u = self._uniform(stream, cols, self._ids(frame, var)).reshape(-1)
        assert tuple(u.shape) == tuple(size), (line, u.shape, size)
        return u

    def normal(self, mean, std, size=None, **kw):
        """Stands in for torch.normal(mean, std, size=...) (fpv_asymmetry.py:191,324,403-410,576)."""
        self.calls += 1
        tag, line, frame = _site()
        assert tag == "fpv" and mean == 0
        size = tuple(size)
        if line in ROUND_NORMAL_SITES:
            stream, slot, word, clip, step, var = ROUND_NORMAL_SITES[line]
            assert std == 1
            ids = self._ids(frame, var)
            k = px.round_normal(self._words(stream, slot, step)[ids, word], clip)
            out = torch.from_numpy(k.astype(np.float32)).reshape(-1, 1)
        else:
            first = OBS_NORMAL_SITES[line]
            n = int(np.prod(size[1:])) if len(size) > 1 else 1
            zs = []
            for i in range(first, first + n):
                b = self._words(O, i // 4)
                pair = px.box_muller(b[:, 0], b[:, 1]) if (i % 4) < 2 else px.box_muller(b[:, 2], b[:, 3])
                zs.append(torch.from_numpy(pair[i % 2]))
            out = torch.stack(zs, dim=1).reshape(size) * std
        assert tuple(out.shape) == size, (line, out.shape, size)
        return out


class _TorchProxy:
    """``torch`` as seen by fpv_asymmetry.py: rand / normal come from the active feeder."""

    def __init__(self, holder):
        self._h = holder

    def __getattr__(self, name):
        if name == "rand":
            return self._h["feeder"].rand
        if name == "normal":
            return self._h["feeder"].normal
        return getattr(torch, name)


_holder = {"feeder": None}


def install(fpv_module):
    """Patch the globals of the loaded reference modules once; the active feeder is switched with ``activate``."""
    if getattr(fpv_module, "_taco_draws_installed", False):
        return
    thrust = sys.modules["isaacgymenvs.tasks.control.thrust_dynamics"]
    battery = sys.modules["isaacgymenvs.tasks.control.battery_dynamics"]
    shim = lambda lo, hi, shape, device: _holder["feeder"].rand_float(lo, hi, shape, device)
    for mod in (fpv_module, thrust, battery):
        mod.torch_rand_float = shim
    fpv_module.torch = _TorchProxy(_holder)
    fpv_module._taco_draws_installed = True


def activate(feeder):
    _holder["feeder"] = feeder

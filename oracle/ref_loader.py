"""Import the reference's own torch modules WITHOUT modifying them.  TEST INFRASTRUCTURE,
build-container only: /root/reference does not exist on the GPU box, so nothing in
tests -m gpu, smoke() or bench.py may call this.  Used by oracle/make_golden.py to
produce tests/golden/*.npz and by tests/test_oracle_vs_reference.py (skipped when the
reference tree is absent).

The control / reward / quaternion modules are plain torch; they import once three
things are stubbed:
  * a fake ``isaacgym`` package whose ``torch_utils`` is loaded by path,
  * dummy ``matplotlib`` / ``matplotlib.pyplot`` (plot-only imports),
  * namespace packages ``isaacgymenvs`` / ``isaacgymenvs.utils``.
``compute_rotating_reward`` allocates on 'cuda:0' inside TorchScript
(task_reward.py:61,77); its source text is read and compiled with the device string
replaced at load time -- the file on disk is never touched.
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("TACO_REFERENCE", "/root/reference")
_ENVS = os.path.join(REF, "IsaacGymEnvs", "isaacgymenvs")


def available():
    return os.path.isdir(_ENVS)


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_cache = {}


def load():
    """Returns a namespace with torch_utils, jit_utils, angvel_control, fpv_dynamics,
    battery_dynamics, thrust_dynamics, task_reward (cpu-patched)."""
    if _cache:
        return _cache["ns"]
    if not available():
        raise RuntimeError("reference tree not found at " + REF)
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.use = lambda *a, **k: None
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    ig = types.ModuleType("isaacgym")
    ig.__path__ = []
    sys.modules["isaacgym"] = ig
    tu = _load("isaacgym.torch_utils", os.path.join(REF, "python", "isaacgym", "torch_utils.py"))
    ig.torch_utils = tu
    pk = types.ModuleType("isaacgymenvs"); pk.__path__ = [_ENVS]
    sys.modules["isaacgymenvs"] = pk
    pu = types.ModuleType("isaacgymenvs.utils"); pu.__path__ = [os.path.join(_ENVS, "utils")]
    sys.modules["isaacgymenvs.utils"] = pu
    jit = _load("isaacgymenvs.utils.torch_jit_utils", os.path.join(_ENVS, "utils", "torch_jit_utils.py"))
    ctrl = os.path.join(_ENVS, "tasks", "control")
    ns = types.SimpleNamespace(torch_utils=tu, jit_utils=jit)
    ns.angvel_control = _load("ref_angvel_control", os.path.join(ctrl, "angvel_control.py"))
    ns.fpv_dynamics = _load("ref_fpv_dynamics", os.path.join(ctrl, "fpv_dynamics.py"))
    ns.battery_dynamics = _load("ref_battery_dynamics", os.path.join(ctrl, "battery_dynamics.py"))
    ns.thrust_dynamics = _load("ref_thrust_dynamics", os.path.join(ctrl, "thrust_dynamics.py"))
    # task_reward: compile from text with 'cuda:0' -> 'cpu' (load-time patch, file untouched)
    src_path = os.path.join(ctrl, "task_reward.py")
    with open(src_path, "r", encoding="utf-8") as fh:
        src = fh.read().replace("'cuda:0'", "'cpu'")
    import tempfile
    tmp_dir = tempfile.mkdtemp(prefix="taco_ref_")           # TorchScript needs a file; never written into the repo
    patched = os.path.join(tmp_dir, "task_reward_cpu.py")
    with open(patched, "w", encoding="utf-8") as fh:
        fh.write(src)
    ns.task_reward = _load("ref_task_reward_cpu", patched)
    nets = os.path.join(REF, "IsaacGymEnvs", "algorithms", "nets_asymmetry.py")
    ns.nets = _load("ref_nets_asymmetry", nets)
    # ppo_asymmetry / buffer_asymmetry import each other as `algorithms.*` (ppo_asymmetry.py:16-22, buffer_asymmetry.py:3-6)
    alg = types.ModuleType("algorithms"); alg.__path__ = [os.path.join(REF, "IsaacGymEnvs", "algorithms")]
    sys.modules.setdefault("algorithms", alg)
    sys.modules["algorithms.nets_asymmetry"] = ns.nets
    ns.buffer = _load("algorithms.buffer_asymmetry", os.path.join(REF, "IsaacGymEnvs", "algorithms", "buffer_asymmetry.py"))
    try:
        ns.ppo = _load("algorithms.ppo_asymmetry", os.path.join(REF, "IsaacGymEnvs", "algorithms", "ppo_asymmetry.py"))
    except Exception as exc:            # tensorboard missing etc.: the PPO class is only needed for spectral_normalize_actors
        ns.ppo = None
        ns.ppo_error = exc
    _cache["ns"] = ns
    return ns


def make_allocator(ns):
    """FpvDynamicsReal2Sim.__init__ puts its weight on cuda:0 (fpv_dynamics.py:28-33);
    build the object without __init__ and give it the same 4x4 constant on CPU."""
    import torch
    obj = object.__new__(ns.fpv_dynamics.FpvDynamicsReal2Sim)
    obj.weight = torch.tensor([[1, -1, 1, -1], [1, -1, -1, 1], [1, 1, -1, -1], [1, 1, 1, 1]], dtype=torch.float32)
    return obj

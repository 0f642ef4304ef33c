"""A fake ``isaacgym`` (gymapi / gymtorch / gymutil) and ``gym.spaces`` that are just big enough to
instantiate the REFERENCE'S OWN env classes -- FpvPos / FpvRotate / FpvFlip / FpvMix
(IsaacGymEnvs/isaacgymenvs/tasks/fpv_asymmetry.py) on VecTask (tasks/base/vec_task_asymmetry.py) --
on the CPU, unmodified.  TEST INFRASTRUCTURE, build container only (needs /root/reference).

Why: the reference env cannot run anywhere (closed IsaacGym / PhysX binaries are absent), so the
``FpvBase`` glue restated in oracle/fpv_env.py had nothing to be pinned against.  With this fake the reference
classes execute every line of their own pre/mid/post_physics_step, refresh_state, compute_observation_state,
reset_idx, reset_*_idx and command logic; oracle/make_golden.py records the trajectories into
tests/golden/glue_*.npz and tests/test_oracle_glue.py holds RefFpvEnv to them.

What the fake provides (every gymapi symbol the two reference files touch: fpv_asymmetry.py:25,124-127,203-310,
335-336,508,633-635; vec_task_asymmetry.py:21,166-179,194,313,317,420-451):
  * ``gymapi.acquire_gym()`` -> FakeGym; SimParams (``dt`` is a C float: reads back as float32), Vec3, Quat,
    Transform, PlaneParams, AssetOptions, enum constants;
  * the actor root-state tensor: one torch CPU tensor (2N, 13), actor 2i = MAV of env i, actor 2i+1 = target marker
    (creation order at fpv_asymmetry.py:275,302), default pose (0,0,4) identity at rest (:264-266);
  * ``gymtorch.wrap_tensor / unwrap_tensor`` = identity (the real ones are zero-copy views, gymtorch.py:61-106);
  * ``apply_rigid_body_force_tensors(LOCAL_SPACE)`` latches the (N,10,3) body force / torque tensors;
    ``simulate`` advances every MAV by oracle.rigid_body (OUR integrator specification -- PhysX is absent) with
    the net body-frame wrench of the latched tensors: rotor bodies 2,4,6,8 at the hub positions of
    assets/xml/fpv_without_duct.xml (parsed in ``load_asset`` and checked against oracle.rigid_body's constants),
    chassis body 0.  The simulator state IS the root tensor ("a reset is visible immediately"):
    each ``simulate`` converts the world-frame angular velocity to the body frame, integrates, converts back.
  * everything graphical is a no-op.

Load-time patches (files on disk are never touched): ``'cuda:0'`` -> ``'cpu'`` in control/fpv_dynamics.py:33 and
control/task_reward.py:61,77 (the trick ref_loader already uses); ``numpy.Inf`` (removed in numpy 2, used at
vec_task_asymmetry.py:94-100) is aliased to ``numpy.inf``.
"""
import importlib
import os
import sys
import tempfile
import types
import xml.etree.ElementTree as ET

import numpy as np
import torch

from . import ref_loader
from . import rigid_body as rb
from .leaf_math import qconj, qrot


# ----------------------------------------------------------------------------------------------- gymapi value types
class Vec3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = float(x), float(y), float(z)


class Quat:
    def __init__(self, x=0.0, y=0.0, z=0.0, w=1.0):
        self.x, self.y, self.z, self.w = float(x), float(y), float(z), float(w)


class Transform:
    def __init__(self):
        self.p, self.r = Vec3(), Quat()


class _Bag:
    """Attribute bag (PlaneParams, AssetOptions, PhysXParams, CameraProperties ...)."""


class SimParams:
    """``dt`` is a C ``float`` member of the real struct: python reads back float32(dt) (fpv_asymmetry.py:219)."""

    def __init__(self):
        self._dt = float(np.float32(1.0 / 60.0))
        self.substeps = 2
        self.num_client_threads = 0
        self.use_gpu_pipeline = False
        self.up_axis = 1
        self.gravity = Vec3(0.0, 0.0, -9.8)
        self.physx = _Bag()
        self.flex = _Bag()

    @property
    def dt(self):
        return self._dt

    @dt.setter
    def dt(self, v):
        self._dt = float(np.float32(v))


class _Sim:
    def __init__(self, params):
        self.dt = params.dt
        self.substeps = int(params.substeps)           # copied at creation: fpv_asymmetry.py:220 sets it afterwards, no effect
        self.gravity_z = params.gravity.z
        self.actor_poses = []                          # (env index, Transform) in creation order
        self.root = None
        self.forces = None
        self.torques = None
        self.n_simulate = 0


class FakeGym:
    """The ``gym`` object returned by ``gymapi.acquire_gym()``."""

    def __init__(self):
        self.asset = None

    # -- construction (vec_task_asymmetry.py:193-194, fpv_asymmetry.py:213-312)
    def create_sim(self, compute_device, graphics_device, physics_engine, sim_params):
        return _Sim(sim_params)

    def add_ground(self, sim, plane_params):
        pass

    def load_asset(self, sim, root, file, options):
        """Parses the reference's MJCF and checks the constants our integrator specification uses."""
        tree = ET.parse(os.path.join(root, file))
        chassis = tree.getroot().find("worldbody").find("body")
        inert = chassis.find("inertial")
        mass = float(inert.get("mass"))
        diag = tuple(float(v) for v in inert.get("diaginertia").split())
        arms, light = [], 0.0
        for arm in chassis.findall("body"):
            ax, ay, az = (float(v) for v in arm.get("pos").split())
            rotor = arm.find("body")
            rz = float(rotor.get("pos").split()[2])
            arms.append((ax, ay, az + rz))
            light += float(arm.find("inertial").get("mass")) + float(rotor.find("inertial").get("mass"))
        assert abs(mass + light - rb.MASS) < 1e-12 and diag == rb.INERTIA, (mass, light, diag)
        # sim rotor bodies 2,4,6,8 = hubs (+x+y, -x+y, -x-y, +x-y)
        assert arms == [(rb.ARM_X, rb.ARM_Y, 0.02), (-rb.ARM_X, rb.ARM_Y, 0.02), (-rb.ARM_X, -rb.ARM_Y, 0.02),
                        (rb.ARM_X, -rb.ARM_Y, 0.02)], arms
        assert options.linear_damping == 0.0 and options.angular_damping == 0.0 and not options.fix_base_link
        self.asset = dict(mass=mass + light, inertia=diag, arms=arms, bodies=1 + 2 * len(arms))
        return self.asset

    def create_sphere(self, sim, radius, options):
        return dict(bodies=1, fixed=bool(options.fix_base_link))

    def create_env(self, sim, lower, upper, num_per_row):
        return len(sim.actor_poses) // 2

    def create_actor(self, env, asset, pose, name, group, filt, seg=0):
        self._sim_of_env.actor_poses.append((env, pose))
        return len(self._sim_of_env.actor_poses) - 1

    def get_actor_dof_properties(self, env, handle):
        return {k: np.zeros(0, dtype=np.float32) for k in ("driveMode", "stiffness", "damping")}   # the MJCF has no joints

    def set_actor_dof_properties(self, env, handle, props):
        pass

    def set_rigid_body_color(self, *a):
        pass

    def get_env_origin(self, env):
        return Vec3()

    def prepare_sim(self, sim):
        n = len(sim.actor_poses)
        root = torch.zeros(n, 13, dtype=torch.float32)
        for i, (_, pose) in enumerate(sim.actor_poses):
            root[i, 0:3] = torch.tensor([pose.p.x, pose.p.y, pose.p.z])
            root[i, 3:7] = torch.tensor([pose.r.x, pose.r.y, pose.r.z, pose.r.w])
        sim.root = root
        return True

    # -- tensors (fpv_asymmetry.py:124-127,335-336,508,633-635)
    def acquire_actor_root_state_tensor(self, sim):
        return sim.root

    def acquire_dof_state_tensor(self, sim):
        return torch.zeros(0, 2)

    def refresh_actor_root_state_tensor(self, sim):
        return True                                    # the tensor is the state

    def refresh_dof_state_tensor(self, sim):
        return True

    def set_actor_root_state_tensor_indexed(self, sim, root_tensor, indices, count):
        assert root_tensor is sim.root and len(indices) == count
        return True                                    # already written in place by the env: visible immediately

    def apply_rigid_body_force_tensors(self, sim, forces, torques, space):
        assert space == LOCAL_SPACE
        sim.forces, sim.torques = forces.clone(), torques.clone()
        return True

    # -- the physics step (vec_task_asymmetry.py:313)
    def simulate(self, sim):
        n = sim.root.shape[0] // 2
        mav = sim.root.view(n, 2, 13)[:, 0, :]
        if sim.forces is None:
            f = torch.zeros(n, 10, 3)
            t = torch.zeros(n, 10, 3)
        else:
            f, t = sim.forces, sim.torques
        rot = [2, 4, 6, 8]
        # only what fpv_asymmetry.py:620-627 can write may be non-zero
        assert float(f[:, rot, 0:2].abs().max()) == 0 and float(t[:, rot, 0:2].abs().max()) == 0
        assert float(f[:, [1, 3, 5, 7, 9]].abs().max()) == 0 and float(t[:, [0, 1, 3, 5, 7, 9]].abs().max()) == 0
        force_b, torque_b = rb.body_wrench(f[:, rot, 2], t[:, rot, 2], f[:, 0, :])
        q = mav[:, 3:7].clone()
        w_b = qrot(qconj(q), mav[:, 10:13].clone())
        pos, quat, vel, w_b = rb.integrate(mav[:, 0:3].clone(), q, mav[:, 7:10].clone(), w_b, force_b, torque_b,
                                           sim.dt, sim.substeps)
        mav[:, 0:3] = pos
        mav[:, 3:7] = quat
        mav[:, 7:10] = vel
        mav[:, 10:13] = qrot(quat, w_b)
        sim.forces = sim.torques = None                # applied forces last one step
        sim.n_simulate += 1

    def fetch_results(self, sim, wait):
        pass

    # -- viewer: never created (headless=True)
    def create_viewer(self, *a):
        raise RuntimeError("fake gym is headless")


def _acquire_gym():
    g = FakeGym()
    orig = g.create_sim

    def create_sim(*a, **k):                           # create_actor gets an env handle, not the sim: remember it
        s = orig(*a, **k)
        g._sim_of_env = s
        return s
    g.create_sim = create_sim
    return g


SIM_PHYSX, SIM_FLEX = 0, 1
UP_AXIS_Y, UP_AXIS_Z = 0, 1
LOCAL_SPACE, ENV_SPACE, GLOBAL_SPACE = 0, 1, 2
DOF_MODE_POS = 1
MESH_VISUAL, MESH_VISUAL_AND_COLLISION = 1, 3


def _install_fake_modules():
    gymapi = types.ModuleType("isaacgym.gymapi")
    for k, v in dict(acquire_gym=_acquire_gym, SimParams=SimParams, Vec3=Vec3, Quat=Quat, Transform=Transform,
                     PlaneParams=_Bag, AssetOptions=_Bag, CameraProperties=_Bag, ContactCollection=int,
                     SIM_PHYSX=SIM_PHYSX, SIM_FLEX=SIM_FLEX, UP_AXIS_Y=UP_AXIS_Y, UP_AXIS_Z=UP_AXIS_Z,
                     LOCAL_SPACE=LOCAL_SPACE, ENV_SPACE=ENV_SPACE, GLOBAL_SPACE=GLOBAL_SPACE, DOF_MODE_POS=DOF_MODE_POS,
                     MESH_VISUAL=MESH_VISUAL, MESH_VISUAL_AND_COLLISION=MESH_VISUAL_AND_COLLISION,
                     KEY_ESCAPE=0, KEY_V=1).items():
        setattr(gymapi, k, v)
    gymtorch = types.ModuleType("isaacgym.gymtorch")
    gymtorch.wrap_tensor = lambda t: t                 # gymtorch.py:61-94: zero-copy view of simulator memory
    gymtorch.unwrap_tensor = lambda t: t               # gymtorch.py:97-106: raw pointer of a contiguous tensor
    gymutil = types.ModuleType("isaacgym.gymutil")
    ig = sys.modules["isaacgym"]
    for name, mod in (("gymapi", gymapi), ("gymtorch", gymtorch), ("gymutil", gymutil)):
        sys.modules["isaacgym." + name] = mod
        setattr(ig, name, mod)
    # gym.spaces.Box (vec_task_asymmetry.py:18-19,94-96); gym itself is not installed here
    gym = types.ModuleType("gym")
    spaces = types.ModuleType("gym.spaces")

    class Box:
        def __init__(self, low, high):
            self.low, self.high = np.asarray(low, dtype=np.float32), np.asarray(high, dtype=np.float32)
            self.shape = self.low.shape
    spaces.Box = Box
    gym.spaces, gym.Space = spaces, object
    sys.modules.setdefault("gym", gym)
    sys.modules.setdefault("gym.spaces", spaces)
    if not hasattr(np, "Inf"):
        np.Inf = np.inf                                # numpy >= 2 dropped the alias the reference uses


def _load_patched(name, path):
    """Import ``path`` under ``name`` with every 'cuda:0' replaced by 'cpu' (through a temp copy)."""
    with open(path, "r", encoding="utf-8") as fh:
        src = fh.read().replace("'cuda:0'", "'cpu'")
    tmp = os.path.join(tempfile.mkdtemp(prefix="taco_ref_"), os.path.basename(path))
    with open(tmp, "w", encoding="utf-8") as fh:
        fh.write(src)
    mod = ref_loader._load(name, tmp)
    mod.__reference_file__ = path
    return mod


_cache = {}


def load_env_module():
    """Returns the reference's ``isaacgymenvs.tasks.fpv_asymmetry`` module, importable on this CPU box."""
    if _cache:
        return _cache["fpv"]
    ref_loader.load()                                  # isaacgym.torch_utils, matplotlib stubs, isaacgymenvs namespaces
    _install_fake_modules()
    envs = ref_loader._ENVS
    for pkg, sub in (("isaacgymenvs.tasks", "tasks"), ("isaacgymenvs.tasks.base", "tasks/base"),
                     ("isaacgymenvs.tasks.control", "tasks/control")):
        m = types.ModuleType(pkg)
        m.__path__ = [os.path.join(envs, sub)]
        m.__package__ = pkg
        sys.modules[pkg] = m
    ctrl = os.path.join(envs, "tasks", "control")
    _load_patched("isaacgymenvs.tasks.control.fpv_dynamics", os.path.join(ctrl, "fpv_dynamics.py"))
    _load_patched("isaacgymenvs.tasks.control.task_reward", os.path.join(ctrl, "task_reward.py"))
    fpv = importlib.import_module("isaacgymenvs.tasks.fpv_asymmetry")
    _cache["fpv"] = fpv
    return fpv


def make_env(task, cfg, cls=None):
    """Instantiate the reference class for ``task`` ('pos' | 'rotate' | 'flip' | 'mix') on the CPU."""
    fpv = load_env_module()
    vt = sys.modules["isaacgymenvs.tasks.base.vec_task_asymmetry"]
    vt.EXISTING_SIM = None                             # vec_task_asymmetry.py:39-45 caches one sim per process
    cls = cls or {"pos": fpv.FpvPos, "rotate": fpv.FpvRotate, "flip": fpv.FpvFlip, "mix": fpv.FpvMix}[task]
    return cls(cfg, "cpu", "cpu", -1, True, False, False)

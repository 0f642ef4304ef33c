"""Host-side mirror of the actor branch of ``PPO_ActorCritic`` for rollout inference.

Mirrors (IsaacGymEnvs/algorithms/):
  * ``MLP.forward``                        nets_asymmetry.py:23-39
  * ``PPO_ActorCritic.act`` (actor branch) nets_asymmetry.py:326-346
  * ``PPO.spectral_normalize_actors``      ppo_asymmetry.py:398-404  (``load(..., lipschitz_const=c)``)

The math runs in libtaco_b200.so (FP32 CUDA-core kernel, or the tcgen05 bf16 kernel with
``tensor_cores=True``); torch only provides device memory and the stream.  No CPU fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _capi


def _as_f32(x, device_id):
    """Contiguous float32 view for the loaders: CUDA tensors on the kernels' device are passed as they are (device pointer, no host
    round trip); anything else becomes a numpy array."""
    if isinstance(x, torch.Tensor):
        x = x.detach()
        if x.device.type == "cuda" and (x.device.index or 0) == device_id:
            return x.to(torch.float32).contiguous()
        return np.ascontiguousarray(x.cpu().numpy(), dtype=np.float32)
    return np.ascontiguousarray(x, dtype=np.float32)


def _ptr(x):
    return x.data_ptr() if isinstance(x, torch.Tensor) else x.ctypes.data


class ActorMLP:
    def __init__(self, input_size, hidden_size, output_size=_capi.NUM_ACTS, device="cuda:0"):
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA is not available: the actor kernels have no CPU fallback")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError(f"ActorMLP runs on CUDA only (got {device!r})")
        self.device_id = dev.index if dev.index is not None else 0
        self.device = f"cuda:{self.device_id}"
        self.sizes = [int(input_size)] + [int(h) for h in hidden_size] + [int(output_size)]
        self._lib = _capi.lib()
        arr = (C.c_int32 * len(self.sizes))(*self.sizes)
        h = C.c_void_p()
        _capi.check(self._lib.taco_actor_create(self.device_id, arr, len(self.sizes), C.byref(h)), "taco_actor_create")
        self._h = h
        self.log_std = np.zeros(self.sizes[-1], dtype=np.float32)      # nets_asymmetry.py:315: log(1.0) * ones

    @property
    def n_layers(self):
        return len(self.sizes) - 1

    @property
    def tensor_cores_available(self):
        return bool(self._lib.taco_actor_tc_available(self._h))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device_id).cuda_stream)

    def load(self, weights, biases, lipschitz_const=0.0, log_std=None):
        """weights[l]: (out, in) like nn.Linear.weight; biases[l]: (out,).  torch tensors or numpy arrays.
        lipschitz_const > 0 applies the spectral projection on the device, once (call after every update)."""
        if len(weights) != self.n_layers or len(biases) != self.n_layers:
            raise ValueError(f"expected {self.n_layers} layers")
        ws, bs = [], []
        for l in range(self.n_layers):
            w, b = _as_f32(weights[l], self.device_id), _as_f32(biases[l], self.device_id)
            if tuple(w.shape) != (self.sizes[l + 1], self.sizes[l]) or tuple(b.shape) != (self.sizes[l + 1],):
                raise ValueError(f"layer {l}: expected weight {(self.sizes[l + 1], self.sizes[l])}, bias {(self.sizes[l + 1],)}")
            ws.append(w); bs.append(b)
        wp = (C.c_void_p * self.n_layers)(*[_ptr(w) for w in ws])
        bp = (C.c_void_p * self.n_layers)(*[_ptr(b) for b in bs])
        _capi.check(self._lib.taco_actor_load(self._h, wp, bp, float(lipschitz_const), self._stream()), "taco_actor_load")
        if log_std is not None:
            self.set_log_std(log_std)

    def set_log_std(self, log_std):
        """The policy's ``log_std`` parameter (nets_asymmetry.py:315,338).  The kernels read exp(log_std)^2 and the log-prob constant
        from device memory owned by the actor, so this also takes effect in launches already captured in a CUDA graph."""
        v = np.ascontiguousarray(log_std.detach().cpu().numpy() if isinstance(log_std, torch.Tensor) else log_std, dtype=np.float32).reshape(-1)
        if v.shape != (self.sizes[-1],):
            raise ValueError(f"log_std must have {self.sizes[-1]} entries")
        self.log_std = v
        _capi.check(self._lib.taco_actor_set_log_std(self._h, v.ctypes.data_as(C.c_void_p), self._stream()), "taco_actor_set_log_std")

    def load_module(self, actor_mlp, log_std=None, lipschitz_const=0.0):
        """Load from a reference-style ``MLP`` (its ``.layers`` Sequential of Linear / activation modules)."""
        lin = [m for m in actor_mlp.layers if isinstance(m, torch.nn.Linear)]
        self.load([m.weight for m in lin], [m.bias for m in lin], lipschitz_const, log_std)

    def sigmas(self):
        out = np.zeros(self.n_layers, dtype=np.float64)
        _capi.check(self._lib.taco_actor_sigmas(self._h, out.ctypes.data_as(C.c_void_p)), "taco_actor_sigmas")
        return out

    def weights(self, layer):
        w = np.empty((self.sizes[layer + 1], self.sizes[layer]), dtype=np.float32)
        b = np.empty(self.sizes[layer + 1], dtype=np.float32)
        _capi.check(self._lib.taco_actor_weights(self._h, layer, w.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p)), "taco_actor_weights")
        return w, b

    def _check_obs(self, obs):
        if obs.device.type != "cuda" or (obs.device.index or 0) != self.device_id:
            obs = obs.to(self.device)
        obs = obs.float().contiguous().view(obs.size(0), -1)          # nets_asymmetry.py:38
        if obs.size(1) != self.sizes[0]:
            raise ValueError(f"actor input must flatten to {self.sizes[0]} features, got {obs.size(1)}")
        return obs

    def forward(self, obs, tensor_cores=False, out=None):
        """mean = tanh(MLP(obs)); obs (N, len_obs, 26) or (N, in) on the actor's device."""
        obs = self._check_obs(obs)
        n = obs.size(0)
        if out is None:
            out = torch.empty(n, self.sizes[-1], dtype=torch.float32, device=self.device)
        _capi.check(self._lib.taco_actor_forward(self._h, C.c_void_p(obs.data_ptr()), C.c_void_p(out.data_ptr()), n,
                                                 1 if tensor_cores else 0, self._stream()), "taco_actor_forward")
        return out

    def act(self, obs, step_index, seed=0, env_offset=0, tensor_cores=False, out=None, step_base=0):
        """Returns (action, clipped_action, log_p, mean): the sample of nets_asymmetry.py:336-346 with Philox noise and
        the clip of ppo_asymmetry.py:310.  ``out`` = (action, clipped, log_p, mean) writes into caller tensors (e.g. rows of
        a RolloutBuffer): contiguous float32 (N,4) / (N,4) / (N[,1]) / (N,4) on the actor's device.  ``step_base``: device address of
        a uint32 added to ``step_index`` on the device (the env's graph-mode counter, ``FpvVecTask.step_counter()[0]``)."""
        obs = self._check_obs(obs)
        n, k = obs.size(0), self.sizes[-1]
        if out is None:
            mean = torch.empty(n, k, dtype=torch.float32, device=self.device)
            action = torch.empty_like(mean)
            clipped = torch.empty_like(mean)
            logp = torch.empty(n, dtype=torch.float32, device=self.device)
        else:
            action, clipped, logp, mean = out
            for t, cnt in ((action, n * k), (clipped, n * k), (logp, n), (mean, n * k)):
                if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != cnt or t.device.type != "cuda":
                    raise ValueError("act(out=...): tensors must be contiguous float32 CUDA tensors of (N,4), (N,4), (N,), (N,4)")
        _capi.check(self._lib.taco_actor_act_counter(self._h, C.c_void_p(obs.data_ptr()), n, self.log_std.ctypes.data_as(C.c_void_p),
                                                     int(env_offset), int(seed) & 0xFFFFFFFFFFFFFFFF, int(step_index) & 0xFFFFFFFF,
                                                     C.c_void_p(int(step_base)), C.c_void_p(mean.data_ptr()), C.c_void_p(action.data_ptr()),
                                                     C.c_void_p(clipped.data_ptr()), C.c_void_p(logp.data_ptr()), 1 if tensor_cores else 0,
                                                     self._stream()), "taco_actor_act")
        return action, clipped, logp, mean

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.taco_actor_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def spectral_normalize_(params, lipschitz_const):
    """``PPO.spectral_normalize_actors`` (ppo_asymmetry.py:398-404) on the device, in place and without a host sync: every
    parameter with >= 2 dimensions whose spectral norm exceeds ``lipschitz_const`` is scaled back onto the ball.  ``params``: an
    iterable of CUDA float32 tensors (e.g. ``agent.actor_mlp.parameters()``); biases are skipped like in the reference.
    Returns the (n_matrices,) float64 device tensor of the spectral norms found (before scaling)."""
    mats = [p for p in params if p.dim() >= 2]
    if not mats:
        return torch.zeros(0, dtype=torch.float64)
    dev = mats[0].device
    if dev.type != "cuda":
        raise RuntimeError("spectral_normalize_ runs on CUDA tensors only (no CPU fallback)")
    lib = _capi.lib()
    sig = torch.zeros(len(mats), dtype=torch.float64, device=dev)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for i, p in enumerate(mats):
        w = p.data
        if w.dtype != torch.float32 or not w.is_contiguous() or w.device != dev:
            raise ValueError("spectral_normalize_: parameters must be contiguous float32 tensors on one CUDA device")
        rows = w.shape[0]
        _capi.check(lib.taco_spectral_project(dev.index or 0, C.c_void_p(w.data_ptr()), rows, w.numel() // rows, float(lipschitz_const),
                                              C.c_void_p(sig[i:].data_ptr()), stream), "taco_spectral_project")
    return sig

"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11 "Parallel random numbers: as
easy as 1, 2, 3"), numpy restatement used by the oracle.  TEST INFRASTRUCTURE.

The reference draws every random number from torch's global generator
(``torch_rand_float`` -> ``torch.rand``, python/isaacgym/torch_utils.py:216-219;
``torch.normal`` at fpv_asymmetry.py:191,324,403-410,576), which no other
implementation can reproduce.  Oracle and CUDA kernel therefore share this stream:

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (global_env_id, rl_step_index, slot, stream)

so a draw depends only on (seed, env, step, purpose) -- never on which GPU owns the env.
The same constants / slot table are compiled into taco_b200/csrc/philox.cuh.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

# streams (counter word 3)
STREAM_ACTIONS = 0      # synthetic U(-1,1) actions for benchmarks
STREAM_RESET = 1        # slots 0..9, see fpv_env.RESET_SLOTS
STREAM_COMMAND = 2      # slot 0: [0] flip turn count at progress==500, [1] rotate speed
STREAM_DEPLOY = 3       # slot 0: [0] action deploy time 9/10/11 ms
STREAM_OBS_NOISE = 4    # slots 0..2 -> 12 normals, slot 3 -> 3 uniforms
STREAM_ROTOR_NOISE = 5  # slot = control sub-step

# floor(Phi(x) * 2^32) for x = -2.5, -1.5, -0.5, 0.5, 1.5, 2.5 : integer CDF of
# round(N(0,1)); makes clamp(round(normal)) draws exact integers on every platform.
ROUND_NORMAL_CDF = np.array([0x0196F4E5, 0x111A46D8, 0x4EFC50EE, 0xB103AF11, 0xEEE5B927, 0xFE690B1A],
                            dtype=np.uint64)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy arrays (or scalars). Returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64) & MASK
    c1 = np.asarray(c1, dtype=np.uint64) & MASK
    c2 = np.asarray(c2, dtype=np.uint64) & MASK
    c3 = np.asarray(c3, dtype=np.uint64) & MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def draw(seed, env_ids, step, slot, stream):
    """(N,4) uint32 block for each global env id."""
    seed = int(seed)
    r = philox4x32_10(env_ids, step, slot, stream, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(r, axis=-1)


def u01(x):
    """uint32 -> float32 in [0,1) with 24 random bits (torch.rand's float32 resolution)."""
    return (np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def round_normal(x, clip):
    """clamp(round(N(0,1)), -clip, clip) as an exact integer draw (clip in {1,3})."""
    x = np.asarray(x, dtype=np.uint64)
    k = np.full(x.shape, -3, dtype=np.int64)
    for t in ROUND_NORMAL_CDF:
        k = k + (x >= t)
    return np.clip(k, -clip, clip)


def box_muller(xa, xb):
    """Two N(0,1) float32 samples from two uint32 words (precise log/sin/cos)."""
    u1 = np.float32(1.0) - u01(xa)                  # (0,1]
    u2 = u01(xb)
    r = np.sqrt(np.float32(-2.0) * np.log(u1).astype(np.float32)).astype(np.float32)
    a = (np.float32(6.283185307179586) * u2).astype(np.float32)
    return (r * np.cos(a).astype(np.float32)).astype(np.float32), (r * np.sin(a).astype(np.float32)).astype(np.float32)

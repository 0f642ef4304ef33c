// Philox4x32-10 (Salmon et al., SC'11), device side.  Same constants / counter layout as
// oracle/philox.py: key = (seed_lo, seed_hi), counter = (global_env_id, rl_step, slot, stream).
#pragma once
#include <stdint.h>

namespace taco {

enum : uint32_t {
    STREAM_ACTIONS = 0,
    STREAM_RESET = 1,
    STREAM_COMMAND = 2,
    STREAM_DEPLOY = 3,
    STREAM_OBS_NOISE = 4,
    STREAM_ROTOR_NOISE = 5,
};

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// uint32 -> [0,1) float32 with 24 random bits
__device__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * 5.9604644775390625e-08f; }

// clamp(round(N(0,1)), -clip, clip) as an exact integer draw: integer CDF thresholds floor(Phi(k+-.5) * 2^32)
__device__ __forceinline__ int round_normal(uint32_t x, int clip) {
    int k = -3;
    k += (x >= 0x0196F4E5u);
    k += (x >= 0x111A46D8u);
    k += (x >= 0x4EFC50EEu);
    k += (x >= 0xB103AF11u);
    k += (x >= 0xEEE5B927u);
    k += (x >= 0xFE690B1Au);
    return max(-clip, min(clip, k));
}

// Box-Muller, precise libm (parity with numpy float32 within a few ulp)
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, float& z0, float& z1) {
    const float u1 = 1.0f - u01(xa);
    const float u2 = u01(xb);
    const float r = sqrtf(-2.0f * logf(u1));
    const float a = 6.283185307179586f * u2;
    float s, c;
    sincosf(a, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

}  // namespace taco

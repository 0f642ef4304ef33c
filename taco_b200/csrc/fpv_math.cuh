// Quaternion / vector leaf math of the fused step kernel.  Operation order follows the
// reference's torch expressions (python/isaacgym/torch_utils.py = TU,
// isaacgymenvs/utils/torch_jit_utils.py = JIT) so that the -fmad=false build rounds like
// eager torch float32.  Quaternions are xyzw.
#pragma once
#include <cuda_runtime.h>

namespace taco {

struct V3 { float x, y, z; };
struct Q4 { float x, y, z, w; };

constexpr float kPi = 3.14159265358979323846f;
constexpr float kTwoPi = 6.28318530717958647692f;
constexpr float kHalfPi = 1.57079632679489661923f;

__device__ __forceinline__ V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 neg(V3 a) { return v3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// a x b with one fused multiply-add per component (integrator only: our own specification, oracle/rigid_body.py)
__device__ __forceinline__ V3 cross_f(V3 a, V3 b) {
    return v3(__fmaf_rn(a.y, b.z, -(a.z * b.y)), __fmaf_rn(a.z, b.x, -(a.x * b.z)), __fmaf_rn(a.x, b.y, -(a.y * b.x)));
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ Q4 conj(Q4 q) { Q4 r; r.x = -q.x; r.y = -q.y; r.z = -q.z; r.w = q.w; return r; }

// TU:58-68 quat_rotate: v(2w^2-1) + 2w(u x v) + 2u(u.v)
__device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
    const float s = 2.0f * q.w * q.w - 1.0f;
    const V3 u = v3(q.x, q.y, q.z);
    const V3 c = cross(u, v);
    const float d = (u.x * v.x + u.y * v.y) + u.z * v.z;
    V3 r;
    r.x = (v.x * s + c.x * q.w * 2.0f) + u.x * d * 2.0f;
    r.y = (v.y * s + c.y * q.w * 2.0f) + u.y * d * 2.0f;
    r.z = (v.z * s + c.z * q.w * 2.0f) + u.z * d * 2.0f;
    return r;
}

// TU:19-40 quat_mul (8-multiply factored form)
__device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
    const float ww = (a.z + a.x) * (b.x + b.y);
    const float yy = (a.w - a.y) * (b.w + b.z);
    const float zz = (a.w + a.y) * (b.w - b.z);
    const float xx = ww + yy + zz;
    const float qq = 0.5f * (xx + (a.z - a.x) * (b.x - b.y));
    Q4 r;
    r.w = qq - ww + (a.z - a.y) * (b.y - b.z);
    r.x = qq - xx + (a.x + a.w) * (b.x + b.w);
    r.y = qq - yy + (a.w - a.x) * (b.y + b.z);
    r.z = qq - zz + (a.z + a.y) * (b.w - b.x);
    return r;
}

// TU:175-196 get_euler_xyz_v1, roll component (the only one the hot path consumes)
// atan2 for finite arguments in ~23 instructions (libdevice's atan2f: ~43): t = min / max of the magnitudes by a reciprocal and
// one Newton correction, atan(t) on [0, 1] as an odd degree-17 minimax polynomial (8 coefficients in t^2, FMA Horner), then the
// octant fix-ups.  Max error 2.4 ulp over 4 M random arguments (tools/atan2_poly_check.py; numpy's float32 arctan2: 3.3 ulp on
// the same set, libdevice documents 2 ulp) -- the same accuracy class as the function it replaces; atan2(0, 0) = 0 like libm.
// The flip task evaluates the roll angle 11 times per RL step (refresh_state in every control iteration, fpv_asymmetry.py:334-360).
__device__ __forceinline__ float atan2_poly(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(fmaxf(ax, ay), 1e-30f), mn = fminf(ax, ay);
    float rc;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(mx));
    float t = mn * rc;
    t = __fmaf_rn(__fmaf_rn(-mx, t, mn), rc, t);
    const float s = t * t;
    float p = 0.00282363896258175373077393f;
    p = __fmaf_rn(p, s, -0.0159569028764963150024414f);
    p = __fmaf_rn(p, s, 0.0425049886107444763183594f);
    p = __fmaf_rn(p, s, -0.0748900920152664184570312f);
    p = __fmaf_rn(p, s, 0.106347933411598205566406f);
    p = __fmaf_rn(p, s, -0.142027363181114196777344f);
    p = __fmaf_rn(p, s, 0.199926957488059997558594f);
    p = __fmaf_rn(p, s, -0.333331018686294555664062f);
    float r = __fmaf_rn(p * s, t, t);
    if (ay > ax) r = kHalfPi - r;
    if (x < 0.0f) r = kPi - r;
    return copysignf(r, y);
}
__device__ __forceinline__ float roll_of(Q4 q) {
    const float sinr = 2.0f * (q.w * q.x + q.y * q.z);
    const float cosr = q.w * q.w - q.x * q.x - q.y * q.y + q.z * q.z;
#ifdef TACO_ATAN2_LIBM                  // tuning / A-B builds only: libdevice's atan2f (5 % slower flip step, profiles/README.md)
    return atan2f(sinr, cosr);
#else
    return atan2_poly(sinr, cosr);
#endif
}

// sin/cos for |x| <= pi/2 by Horner evaluation of the degree-13/14 Taylor polynomials (oracle/leaf_math.py
// sincos_draw): + and * only, so the -fmad=false build matches the torch oracle bit for bit.  Used for the
// random attitude draws of a reset, which are our own Philox-driven draws.
__device__ __forceinline__ void sincos_draw(float x, float* s, float* c) {
    const float z = x * x;
    float a = (float)(1.0 / 6227020800);
    a = (float)(-1.0 / 39916800) + z * a;
    a = (float)(1.0 / 362880) + z * a;
    a = (float)(-1.0 / 5040) + z * a;
    a = (float)(1.0 / 120) + z * a;
    a = (float)(-1.0 / 6) + z * a;
    *s = x * (1.0f + z * a);
    float b = (float)(-1.0 / 87178291200);
    b = (float)(1.0 / 479001600) + z * b;
    b = (float)(-1.0 / 3628800) + z * b;
    b = (float)(1.0 / 40320) + z * b;
    b = (float)(-1.0 / 720) + z * b;
    b = (float)(1.0 / 24) + z * b;
    b = (float)(-1.0 / 2) + z * b;
    *c = 1.0f + z * b;
}

// TU:199-213 quat_from_euler_xyz with sincos_draw for the half angles
__device__ __forceinline__ Q4 quat_from_euler(float roll, float pitch, float yaw) {
    float sy, cy, sr, cr, sp, cp;
    sincos_draw(yaw * 0.5f, &sy, &cy);
    sincos_draw(roll * 0.5f, &sr, &cr);
    sincos_draw(pitch * 0.5f, &sp, &cp);
    Q4 q;
    q.w = cy * cr * cp + sy * sr * sp;
    q.x = cy * sr * cp - sy * cr * sp;
    q.y = cy * cr * sp + sy * sr * cp;
    q.z = sy * cr * cp - cy * sr * sp;
    return q;
}

// JIT:389-416 quaternion_to_matrix, row-major m[9]
__device__ __forceinline__ void rotmat9(Q4 q, float* m) {
    const float i = q.x, j = q.y, k = q.z, r = q.w;
    const float two_s = 2.0f / (((i * i + j * j) + k * k) + r * r);
    m[0] = 1.0f - two_s * (j * j + k * k);
    m[1] = two_s * (i * j - k * r);
    m[2] = two_s * (i * k + j * r);
    m[3] = two_s * (i * j + k * r);
    m[4] = 1.0f - two_s * (i * i + k * k);
    m[5] = two_s * (j * k - i * r);
    m[6] = two_s * (i * k - j * r);
    m[7] = two_s * (j * k + i * r);
    m[8] = 1.0f - two_s * (i * i + j * j);
}

// torch.norm(p=2) on CPU accumulates acc = fma(x_i, x_i, acc) in order and takes one sqrt (verified against
// torch 2.11 in tests/test_oracle_semantics.py); the reference's norms must be formed the same way in BOTH builds.
__device__ __forceinline__ float norm2(float x, float y) { return sqrtf(__fmaf_rn(y, y, x * x)); }
__device__ __forceinline__ float norm3(float x, float y, float z) { return sqrtf(__fmaf_rn(z, z, __fmaf_rn(y, y, x * x))); }

// x / C for a compile-time constant C, correctly rounded (== IEEE x / C) with three FMA-class instructions instead of
// the ~12-instruction IEEE division sequence: q = RN(x*rc), r = x - C*q (exact, one FMA), q' = RN(q + r*rc), rc = RN(1/C)
// (Markstein 1990; valid while x/C neither overflows nor underflows and C's significand is not all ones -- checked
// exhaustively on the GPU by taco_selftest_divc / tests/test_kernel_selftest_gpu.py).  Non-finite x gives NaN.
#define TACO_DIVC(x, C) ::taco::divc_impl((x), (C), 1.0f / (C))
__device__ __forceinline__ float divc_impl(float x, float c, float rc) {
    const float q = x * rc;
    const float r = __fmaf_rn(-c, q, x);
    return __fmaf_rn(r, rc, q);
}

// (hi-lo)*u + lo, TU:216-219 torch_rand_float; rng/lo are the float32 casts of the python doubles
__device__ __forceinline__ float rr(float rng, float lo, float u) { return rng * u + lo; }

// 1/(1+d^2) + 1/(1+10 d^2), task_reward.py:26-28 (note (10*d)*d)
__device__ __forceinline__ float two_scale(float d) { return 1.0f / (1.0f + d * d) + 1.0f / (1.0f + 10.0f * d * d); }

}  // namespace taco

// General tcgen05 GEMM of the native PPO update (taco_ppo.cu):  D[M, N] = A[M, K] * B[N, K]^T, bf16 operands, fp32 accumulation
// in tensor memory.  Both operands are row-major with K contiguous ("K-major") and reach shared memory through TMA tensor maps
// (cp.async.bulk.tensor.2d, SWIZZLE_128B, out-of-bounds boxes zero-filled, so ragged M / N / K need no padding code).
//
//   tile       128 (M) x n_tile (N <= 256, multiple of 16) per work item; K in blocks of 64 (one 128-byte swizzle row), 16 per MMA
//   work item  (m_tile, k_split): split-K gives the K = batch GEMMs of the weight gradients (2 output tiles) enough CTAs; every
//              split writes its own fp32 partial, which the gradient-norm kernel sums in a fixed order (bit-reproducible)
//   pipeline   persistent CTAs; warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2..5 = epilogue;
//              4 smem stages (full / empty mbarriers), 2 accumulators of 256 TMEM columns (tmem_full / tmem_empty), so the
//              epilogue of one work item overlaps the MMAs of the next
//   epilogues  EPI_F32            out_f32[split][m][n] = acc (+ bias[n])
//              EPI_BIAS_RELU_DUAL y = relu(acc + bias[n]) as bf16, written batch-major [m][n] AND feature-major [n][m]
//              EPI_RELUBWD_DUAL   g = acc * (act[m][n] > 0) as bf16, both layouts (the ReLU backward of the layer below)
//              EPI_TANH_F32       out_f32[m][n] = tanh(acc + bias[n])     (actor mean, nets_asymmetry.py:32-39)
//              EPI_LSTM           one LSTM time step from the gate pre-activations (nets_asymmetry.py:128-136), see LstmEpi
// A thread of an epilogue warp owns one accumulator row (= TMEM lane).  A layout with "samples contiguous" is what the next
// GEMM with K = batch needs as its K-major operand; "features contiguous" is what a GEMM with K = features needs: producing
// both in the epilogue keeps every operand of every GEMM K-major without transpose passes.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

#ifndef TACO_TC_NO_ACTOR_KERNEL
#define TACO_TC_NO_ACTOR_KERNEL          // only the PTX wrappers of actor_tc.cuh are used here
#endif
#include "actor_tc.cuh"

namespace taco {
namespace gemm {

using namespace taco::actor;

constexpr int BM = 128, BK = 64;
constexpr int kStages = 4;
constexpr int kGemmThreads = 192;
constexpr int kStageA = BM * BK * 2;          // 16 KB
constexpr int kStageB = 256 * BK * 2;         // 32 KB (n_tile <= 256)
constexpr int kGemmSmem = 1024 + kStages * (kStageA + kStageB) + 256;
constexpr int kAccCols = 256;

enum : int { EPI_F32 = 0, EPI_BIAS_RELU_DUAL = 1, EPI_RELUBWD_DUAL = 2, EPI_TANH_F32 = 3, EPI_LSTM = 4 };

struct LstmEpi {
    // gates = [i | f | g | o] pre-activations of H = 64 units each (torch nn.LSTM row blocks), bias already inside the GEMM
    const float* c_prev;              // [m][64] fp32 (null: zero)
    float* c_out;                     // [m][64] fp32
    __nv_bfloat16* gates_out;         // [m][256] bf16: sigmoid(i), sigmoid(f), tanh(g), sigmoid(o)   (saved for the backward pass)
    __nv_bfloat16* h_bm;              // batch-major destination of h: row stride ld_h_bm elements (the next step's operand [h | x | 1 1])
    long long ld_h_bm;
    __nv_bfloat16* h_fm;              // feature-major destination of h: row j at h_fm + j * ld_h_fm, sample m at + m
    long long ld_h_fm;
};

struct GemmParams {
    int m, n, k;                      // rows of A, rows of B, depth
    int n_tile;                       // UMMA N: multiple of 16, >= n, <= 256
    int splits, kb_per_split;         // split-K: work item (m_tile, split) covers k blocks [split * kb_per_split, ...)
    int a_row0, b_row0;               // row offsets into the tensor maps (views of larger buffers)
    int epi;
    int n_valid;                      // columns written by the epilogue (<= n_tile)
    float* out_f32; long long ldc, split_stride;
    __nv_bfloat16* out_bm; long long ld_bm;
    __nv_bfloat16* out_fm; long long ld_fm;
    const float* bias;
    const __nv_bfloat16* act; long long ld_act;
    const int* stop;                  // device flag (may be null): non-zero = the update stopped early (KL), the launch is a no-op
    LstmEpi lstm;
};

__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst_smem),
                 "l"(tm), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* dst, const float* v) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]); u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(dst) = u;
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_a = smem_u32(smem), s_b = s_a + kStages * kStageA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kStageA + kStageB));
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kStages, bar_tfull = bar_empty + 8 * kStages, bar_tempty = bar_tfull + 16;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(bar_tfull + 8 * b, 1); mbar_init(bar_tempty + 8 * b, 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(s_tmem), 2 * kAccCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *s_tmem;

    const int m_tiles = (p.m + BM - 1) / BM;
    const int total = (p.stop != nullptr && __ldg(p.stop) != 0) ? 0 : m_tiles * p.splits;
    const int kb_total = (p.k + BK - 1) / BK;
    const uint32_t stage_bytes = (uint32_t)(kStageA + p.n_tile * BK * 2);

    if (warp == 0) {
        // ===================== TMA producer
        uint32_t stage = 0, phase = 0;
        for (int w = blockIdx.x; w < total; w += gridDim.x) {
            const int mt = w % m_tiles, sp = w / m_tiles;
            const int kb0 = sp * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, kb_total);
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1u);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(bar_full + 8 * stage, stage_bytes);
                    tma_load_2d(s_a + stage * kStageA, &tmA, kb * BK, p.a_row0 + mt * BM, bar_full + 8 * stage);
                    tma_load_2d(s_b + stage * kStageB, &tmB, kb * BK, p.b_row0, bar_full + 8 * stage);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer
        uint32_t stage = 0, phase = 0, buf = 0, acc_phase = 0;
        const uint32_t idesc = umma_idesc_bf16(BM, p.n_tile);
        for (int w = blockIdx.x; w < total; w += gridDim.x) {
            const int sp = w / m_tiles;
            const int kb0 = sp * p.kb_per_split, kb1 = min(kb0 + p.kb_per_split, kb_total);
            mbar_wait(bar_tempty + 8 * buf, acc_phase ^ 1u);           // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_addr = tmem0 + buf * kAccCols;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t adesc = umma_desc_sw128(s_a + stage * kStageA), bdesc = umma_desc_sw128(s_b + stage * kStageB);
                    const int ksteps = min(BK / 16, (p.k - kb * BK + 15) / 16);      // the zero-filled tail of K costs no MMAs
                    for (int ks = 0; ks < ksteps; ++ks)
                        umma_bf16(d_addr, adesc + (uint64_t)(2 * ks), bdesc + (uint64_t)(2 * ks), idesc, (uint32_t)(kb != kb0 || ks != 0));
                    umma_commit(bar_empty + 8 * stage);                 // stage free once these MMAs have read it
                    if (kb == kb1 - 1) umma_commit(bar_tfull + 8 * buf);
                }
                __syncwarp();
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
            }
            if (kb1 <= kb0 && elect_one_sync()) umma_commit(bar_tfull + 8 * buf);   // empty split (cannot happen with the host's partition)
            buf ^= 1u;
            if (buf == 0) acc_phase ^= 1u;
        }
    } else {
        // ===================== epilogue: thread <-> accumulator row (TMEM lane)
        const int quad = warp & 3;                                        // a warp may only touch TMEM lanes 32 * (warp % 4) .. + 31
        const int r = (quad << 5) | lane;
        uint32_t buf = 0, acc_phase = 0;
        for (int w = blockIdx.x; w < total; w += gridDim.x) {
            const int mt = w % m_tiles, sp = w / m_tiles;
            const long long m = (long long)mt * BM + r;
            const bool row_ok = m < p.m;
            mbar_wait(bar_tfull + 8 * buf, acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem0 + ((uint32_t)(quad << 5) << 16) + buf * kAccCols;
            if (p.epi == EPI_LSTM) {
                const LstmEpi& L = p.lstm;
#pragma unroll 1
                for (int j0 = 0; j0 < 64; j0 += 16) {
                    uint32_t vi[16], vf[16], vg[16], vo[16];
                    tmem_ld16(t_row + j0, vi); tmem_ld16(t_row + 64 + j0, vf); tmem_ld16(t_row + 128 + j0, vg); tmem_ld16(t_row + 192 + j0, vo);
                    tmem_ld_wait();
                    if (row_ok) {
                        float cp[16], gi[16], gf[16], gg[16], go[16], hh[16], cc[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 c4 = L.c_prev ? *reinterpret_cast<const float4*>(L.c_prev + m * 64 + j0 + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
                            cp[4 * q] = c4.x; cp[4 * q + 1] = c4.y; cp[4 * q + 2] = c4.z; cp[4 * q + 3] = c4.w;
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            gi[j] = sigmoidf_(__uint_as_float(vi[j])); gf[j] = sigmoidf_(__uint_as_float(vf[j]));
                            gg[j] = tanhf(__uint_as_float(vg[j])); go[j] = sigmoidf_(__uint_as_float(vo[j]));
                            cc[j] = gf[j] * cp[j] + gi[j] * gg[j];
                            hh[j] = go[j] * tanhf(cc[j]);
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<float4*>(L.c_out + m * 64 + j0 + 4 * q) = make_float4(cc[4 * q], cc[4 * q + 1], cc[4 * q + 2], cc[4 * q + 3]);
                        __nv_bfloat16* g = L.gates_out + m * 256 + j0;
                        store_bf16x8(g, gi); store_bf16x8(g + 8, gi + 8);
                        store_bf16x8(g + 64, gf); store_bf16x8(g + 72, gf + 8);
                        store_bf16x8(g + 128, gg); store_bf16x8(g + 136, gg + 8);
                        store_bf16x8(g + 192, go); store_bf16x8(g + 200, go + 8);
                        if (L.h_bm) { store_bf16x8(L.h_bm + m * L.ld_h_bm + j0, hh); store_bf16x8(L.h_bm + m * L.ld_h_bm + j0 + 8, hh + 8); }
                        if (L.h_fm) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) L.h_fm[(long long)(j0 + j) * L.ld_h_fm + m] = __float2bfloat16_rn(hh[j]);
                        }
                    }
                }
            } else {
                for (int c0 = 0; c0 < p.n_valid; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(t_row + c0, v);
                    tmem_ld_wait();
                    if (!row_ok) continue;
                    float y[32];
                    const int nv = min(32, p.n_valid - c0);
                    if (p.epi == EPI_F32) {
                        float* o = p.out_f32 + (long long)sp * p.split_stride + m * p.ldc + c0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) y[j] = __uint_as_float(v[j]) + ((p.bias && j < nv) ? __ldg(p.bias + c0 + j) : 0.0f);
                        if (nv == 32 && ((reinterpret_cast<uintptr_t>(o) & 15u) == 0)) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(o + 4 * q) = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
                        } else {
                            for (int j = 0; j < nv; ++j) o[j] = y[j];
                        }
                    } else if (p.epi == EPI_TANH_F32) {
                        float* o = p.out_f32 + m * p.ldc + c0;
                        for (int j = 0; j < nv; ++j) o[j] = tanhf(__uint_as_float(v[j]) + __ldg(p.bias + c0 + j));
                    } else {
                        if (p.epi == EPI_BIAS_RELU_DUAL) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) y[j] = (j < nv) ? fmaxf(__uint_as_float(v[j]) + __ldg(p.bias + c0 + j), 0.0f) : 0.0f;
                        } else {                                      // EPI_RELUBWD_DUAL
                            const __nv_bfloat16* a = p.act + m * p.ld_act + c0;
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint4 u = make_uint4(0u, 0u, 0u, 0u);
                                if (8 * q < nv) u = *reinterpret_cast<const uint4*>(a + 8 * q);
                                const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    // bf16 > 0  <=>  sign clear and not zero (activations are post-ReLU: never negative, never NaN)
                                    const bool lo = (w4[e] & 0x7FFFu) != 0u && !(w4[e] & 0x8000u), hi = (w4[e] & 0x7FFF0000u) != 0u && !(w4[e] & 0x80000000u);
                                    y[8 * q + 2 * e] = lo ? __uint_as_float(v[8 * q + 2 * e]) : 0.0f;
                                    y[8 * q + 2 * e + 1] = hi ? __uint_as_float(v[8 * q + 2 * e + 1]) : 0.0f;
                                }
                            }
                        }
                        if (p.out_bm) {
                            __nv_bfloat16* o = p.out_bm + m * p.ld_bm + c0;
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (8 * q < nv) store_bf16x8(o + 8 * q, y + 8 * q);
                        }
                        if (p.out_fm) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < nv) p.out_fm[(long long)(c0 + j) * p.ld_fm + m] = __float2bfloat16_rn(y[j]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
            buf ^= 1u;
            if (buf == 0) acc_phase ^= 1u;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem0, 2 * kAccCols);
}

}  // namespace gemm
}  // namespace taco

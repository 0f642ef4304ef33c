// Actor MLP on the 5th-generation tensor cores (tcgen05 + TMEM), one persistent CTA per SM.
//
// Replaces, for rollout inference, MLP.forward + the actor branch of PPO_ActorCritic.act
// (IsaacGymEnvs/algorithms/nets_asymmetry.py:23-39, :326-346):  mean = tanh(W_L relu(... relu(W_1 x + b_1) ...) + b_L).
//
// Shape of the work: a tile is 128 envs (rows = TMEM lanes); a CTA keeps TWO tiles in flight (a pair).  Every hidden
// layer is D[128 x N] = A[128 x K] * W[N x K]^T by tcgen05.mma (bf16 operands, fp32 accumulation, M=128, N = one
// 128-row part of W, K=16 per instruction).  The activations never touch shared memory: A lives in TENSOR MEMORY as
// packed bf16 pairs (the "TS" form of the instruction, A from TMEM), D in the neighbouring 128 TMEM columns, and the
// tile's epilogue warps turn D into the next layer's A register-side (tcgen05.ld -> +bias, ReLU -> bf16x2 ->
// tcgen05.st).  Shared memory holds only the weight ring: W streams from L2 through 64 KB slots (one 128-row part of a
// layer = up to four 16 KB K-chunks, K-major SWIZZLE_128B) filled by the bulk async-copy engine (cp.async.bulk + mbarrier
// complete_tx) from images that taco_actor_load pre-swizzled once per update; every slot is consumed by BOTH tiles of the
// pair before it is released (half the L2 traffic per env), and the two tiles alternate part by part, so while the tensor core works
// on one tile the epilogue warps of the other drain its accumulator.  The 4-wide output layer rides the same chain as
// one more (N = 16, zero-padded) MMA part; tanh and the optional Gaussian sampling are CUDA-core work in its epilogue.
//
// TMEM map (512 columns): tile slot t owns columns [256 t, 256 t + 256): A = first 128 (K <= 256 as bf16 pairs),
// D = last 128 (one part of N).
// Warp roles (576 threads): warp 0 = weight producer, warp 1 = TMEM allocator + MMA issuer (both warp-uniform loops
// with one elected lane issuing, so that every operand stays in uniform registers),
// warps 2..17 = epilogue: tile slot t = (warp-2)/8, column half of the part = ((warp-2)/4)&1, TMEM lane quadrant =
// warp%4 (thread <-> TMEM lane <-> env row).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include "philox.cuh"

namespace taco {
namespace actor {

constexpr int kTileM = 128;                    // envs per tile = TMEM lanes
constexpr int kKC = 64;                        // bf16 per K chunk: one 128-byte swizzle row
constexpr int kMaxN = 256;                     // widest hidden layer
constexpr int kPartN = 128;                    // rows of W per stage = columns of D per part
constexpr int kSlots = 3;                      // weight ring: slots of one part (all K chunks)
constexpr int kChunkBytes = kPartN * 128;      // 16 KB: 128 rows x 64 bf16
constexpr int kSlotBytes = kChunkBytes * (kMaxN / kKC);   // 64 KB
constexpr int kMaxHidden = 4;
constexpr int kOutPad = 4;                     // num_acts = 4
constexpr int kOutN = 16;                      // output layer as an MMA part: N padded to the instruction minimum
constexpr int kEpiWarps = 16;
constexpr int kTcThreads = (2 + kEpiWarps) * 32;   // 576
constexpr int kTmemCols = 512;
constexpr int kTmemSlot = 256;                 // columns per tile slot: A (128, packed bf16 pairs) + D (128, fp32)
constexpr int kTmemD = 128;                    // offset of D inside a slot
constexpr uint32_t STREAM_ACTOR = 6;           // Philox stream of the action noise

struct TcLayer {
    int n;               // output width of the layer (multiple of 64, <= 256)
    int kchunks;         // ceil(K / 64)
    uint32_t img_off;    // byte offset of the layer's first chunk image; chunk c is at img_off + c * n * 128
};

struct SampleParams {    // PPO_ActorCritic.act sampling (nets_asymmetry.py:336-346); enabled when action != nullptr
    float* action;       // (n, out) mean + std * eps
    float* clipped;      // (n, out) clamp(action, -1, 1)   (ppo_asymmetry.py:310)
    float* logp;         // (n)
    const float* samp;   // device, owned by the TacoActor: [0..3] std = exp(log_std)^2 (scale_tril = diag(exp(log_std)^2),
                         // nets_asymmetry.py:338), [4] logp_const = -sum(log std) - out/2 * log(2 pi).  In device memory, not by
                         // value, so that a launch captured in a CUDA graph follows taco_actor_set_log_std between replays
    long long env_offset;
    uint32_t seed_lo, seed_hi, step_index;
    const uint32_t* step_base;   // device word added to step_index (null: none), see taco_actor_act_counter
};

struct TcParams {
    const float* obs;    // (n_rows, in_dim) f32
    float* mean;         // (n_rows, out_dim) f32
    int in_dim, out_dim, n_rows, num_tiles, n_hidden;
    int tiles_per_cta;   // 2 (a CTA keeps a pair of tiles in flight) or 1
    int obs_bulk;        // in_dim <= 26, even, 16-byte aligned base: whole tiles of observations are staged by the bulk-copy engine
    int obs_vec2;        // rows of obs are 8-byte aligned (even in_dim, aligned base): observation loads use 8-byte accesses
    const uint8_t* wimg; // pre-swizzled bf16 chunk images of the hidden layers and of the (16-row padded) output layer
    const float* bias;   // [kMaxHidden][kMaxN] f32
    const float* b_out;  // [kOutPad]
    TcLayer layer[kMaxHidden + 1];   // hidden layers, then the output layer (n = kOutN)
    SampleParams sp;
    unsigned long long* dbg;   // optional timeline of CTA 0 (TACO_ACTOR_TIMELINE=<file>): 3 regions of kDbgCap (clock << 8 | code) words
};
constexpr int kDbgCap = 4096;
#define TACO_DBG(region, cnt, code)                                                                         \
    do {                                                                                                    \
        if (p.dbg && blockIdx.x == 0 && (cnt) < kDbgCap)                                                    \
            p.dbg[(region) * kDbgCap + (cnt)++] = ((unsigned long long)clock64() << 8) | (unsigned)(code);  \
    } while (0)

constexpr int kSmemBias = kMaxHidden * kMaxN * 4;          // 4 KB
constexpr int kTcSmemBytes = 1024 /*alignment slack*/ + kSlots * kSlotBytes + kSmemBias + 256;
constexpr int kObsBulkMaxIn = 26;              // widest observation row the bulk-staged path takes (26 x len_obs = 1)
constexpr int kObsStageBytes = kTileM * kObsBulkMaxIn * 4;        // one tile's observation block, fp32, as it lies in global memory
constexpr int kActorSmemBytes = kTcSmemBytes + 2 * kObsStageBytes;

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint32_t elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// bulk async copy global -> shared (the TMA engine, 1-D form): completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: 8-row x 128-byte atoms, consecutive atoms 1024 B apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout=2 [61,64))
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor, kind::f16: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, both K-major, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued tcgen05.mma of this thread arrive (once) on the mbarrier when they have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t <-> lane base + t)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> 32 lanes x 16 consecutive columns (thread t <-> lane base + t)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem, packed bf16 pairs] * B[smem descriptor]^T   (the "TS" form: A operand from tensor memory)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);     // .x (low half) = lo
    return *reinterpret_cast<const uint32_t*>(&v);
}

// byte offset of the 16-byte chunk c16 (0..7) of row r inside a [rows x 64 bf16] SWIZZLE_128B block
__device__ __forceinline__ uint32_t sw128_off(int r, int c16) { return (uint32_t)(r * 128 + ((c16 ^ (r & 7)) << 4)); }

// tanh output + optional sampling, shared by the tensor-core and the FP32 kernels (one thread = one env row)
__device__ __forceinline__ void actor_tail(const float* pre, int out_dim, long long row, float* mean, const SampleParams& sp) {
    float mu[kOutPad];
#pragma unroll
    for (int o = 0; o < kOutPad; ++o) mu[o] = tanhf(pre[o]);
    if (out_dim == 4) {
        *reinterpret_cast<float4*>(mean + row * 4) = make_float4(mu[0], mu[1], mu[2], mu[3]);
    } else {
        for (int o = 0; o < out_dim; ++o) mean[row * out_dim + o] = mu[o];
    }
    if (sp.action) {
        const uint4 r = philox4x32_10((uint32_t)(sp.env_offset + row), sp.step_index + (sp.step_base ? __ldg(sp.step_base) : 0u), 0, STREAM_ACTOR,
                                      sp.seed_lo, sp.seed_hi);
        float z[4];
        box_muller(r.x, r.y, z[0], z[1]);
        box_muller(r.z, r.w, z[2], z[3]);
        float lp = __ldg(sp.samp + 4);
        float a[kOutPad] = {0.f, 0.f, 0.f, 0.f};
        for (int o = 0; o < out_dim; ++o) {
            a[o] = mu[o] + __ldg(sp.samp + o) * z[o];
            lp += -0.5f * (z[o] * z[o]);
        }
        if (out_dim == 4 && ((reinterpret_cast<uintptr_t>(sp.action) | reinterpret_cast<uintptr_t>(sp.clipped)) & 15u) == 0) {
            *reinterpret_cast<float4*>(sp.action + row * 4) = make_float4(a[0], a[1], a[2], a[3]);
            if (sp.clipped)
                *reinterpret_cast<float4*>(sp.clipped + row * 4) = make_float4(fminf(fmaxf(a[0], -1.0f), 1.0f), fminf(fmaxf(a[1], -1.0f), 1.0f),
                                                                               fminf(fmaxf(a[2], -1.0f), 1.0f), fminf(fmaxf(a[3], -1.0f), 1.0f));
        } else {
            for (int o = 0; o < out_dim; ++o) {
                sp.action[row * out_dim + o] = a[o];
                if (sp.clipped) sp.clipped[row * out_dim + o] = fminf(fmaxf(a[o], -1.0f), 1.0f);
            }
        }
        if (sp.logp) sp.logp[row] = lp;
    }
}

// ---------------------------------------------------------------------------------------------- the kernel
// relu(D + bias) of 32 accumulator columns -> 16 packed bf16 pairs (column 2j in the low half)
__device__ __forceinline__ void relu_pack32(const uint32_t (&v)[32], const float* bias, uint32_t* pk) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 b = *reinterpret_cast<const float4*>(bias + q * 4);
        pk[q * 2 + 0] = pack_bf16x2(fmaxf(__uint_as_float(v[q * 4 + 0]) + b.x, 0.0f), fmaxf(__uint_as_float(v[q * 4 + 1]) + b.y, 0.0f));
        pk[q * 2 + 1] = pack_bf16x2(fmaxf(__uint_as_float(v[q * 4 + 2]) + b.z, 0.0f), fmaxf(__uint_as_float(v[q * 4 + 3]) + b.w, 0.0f));
    }
}

// the observation features [32 g, 32 g + 32) of one row as 16 packed bf16 pairs (zero beyond in_dim / for invalid rows).  Rows of an
// even width are 8-byte aligned ((N, in_dim) float32 contiguous): one 8-byte load per pair halves the L1 wavefronts of this
// row-per-thread access pattern (every lane touches its own row, i.e. its own cache line).
__device__ __forceinline__ void load_obs_group(const float* x, int g, int in_dim, bool valid, uint32_t* pk, bool vec2) {
    if (vec2) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int k = g * 32 + 2 * j;
            const float2 f = (valid && k < in_dim) ? __ldg(reinterpret_cast<const float2*>(x + k)) : make_float2(0.0f, 0.0f);
            pk[j] = pack_bf16x2(f.x, f.y);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int k = g * 32 + 2 * j;
            const float f0 = (valid && k < in_dim) ? __ldg(x + k) : 0.0f;
            const float f1 = (valid && k + 1 < in_dim) ? __ldg(x + k + 1) : 0.0f;
            pk[j] = pack_bf16x2(f0, f1);
        }
    }
}

// features [0, 32) of row r of a bulk-staged observation block (rows of in_dim floats, as in global memory) as 16 packed bf16
// pairs; 8-byte reads at a row stride of in_dim * 4 bytes (26 words: conflict-free per half-warp)
__device__ __forceinline__ void read_staged_obs(const float* stage, int r, int in_dim, uint32_t* pk) {
    const float2* row = reinterpret_cast<const float2*>(stage + r * in_dim);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const float2 f = (2 * j < in_dim) ? row[j] : make_float2(0.0f, 0.0f);
        pk[j] = pack_bf16x2(f.x, f.y);
    }
}

#ifndef TACO_TC_NO_ACTOR_KERNEL      // critic_tc.cuh reuses the wrappers above and the weight packer below, not this kernel
__global__ void __launch_bounds__(kTcThreads, 1) actor_tc_kernel(const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;                   // SWIZZLE_128B atoms need 1024-byte alignment
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t s_ring = base;                                    // kSlots x 64 KB weight ring
    float* s_bias = reinterpret_cast<float*>(sm + kSlots * kSlotBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + kMaxHidden * kMaxN);
    const uint32_t bar_full = smem_u32(bars);                        // [kSlots] producer -> MMA
    const uint32_t bar_empty = bar_full + 8 * kSlots;                // [kSlots] MMA -> producer
    const uint32_t bar_a = bar_empty + 8 * kSlots;                   // [2] epilogue(t) -> MMA: D(t) drained (and, for a new layer, A(t) written)
    const uint32_t bar_d = bar_a + 16;                               // [2] MMA -> epilogue(t): the part of D(t) is complete
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kSlots + 4);
    const uint32_t bar_obs = bar_d + 16 + 8;                         // [2] a tile's observation block has landed in the staging buffer
    float* s_obs = reinterpret_cast<float*>(sm + kTcSmemBytes - 1024);   // [2 tiles][128 rows x in_dim] (only with p.obs_bulk)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tps = p.tiles_per_cta;                                 // 2, or 1 when there are fewer tiles than SMs (small batches: one tile per CTA halves the latency)
    const int num_pairs = (p.num_tiles + tps - 1) / tps;
    const int n_layers = p.n_hidden + 1;                             // hidden layers + the output layer as a 16-wide part

    for (int i = threadIdx.x; i < kMaxHidden * kMaxN; i += kTcThreads) s_bias[i] = p.bias[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlots; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(bar_a + 8 * t, (kEpiWarps / 2) * 32); mbar_init(bar_d + 8 * t, 1); mbar_init(bar_obs + 8 * t, 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(s_tmem), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *s_tmem;

    if (warp == 0) {
        // ===================== weight producer: per pair, per layer, per 128-row part: one slot (all its K chunks)
        uint32_t slot = 0, phase = 0;
        for (int pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
            for (int l = 0; l < n_layers; ++l) {
                const int n = p.layer[l].n, kch = p.layer[l].kchunks;
                const uint8_t* img = p.wimg + p.layer[l].img_off;
                for (int h0 = 0; h0 < n; h0 += kPartN) {
                    const uint32_t bytes = (uint32_t)min(kPartN, n - h0) * 128u;
                    mbar_wait(bar_empty + 8 * slot, phase ^ 1u);
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(bar_full + 8 * slot, bytes * (uint32_t)kch);
                        for (int c = 0; c < kch; ++c)
                            bulk_g2s(s_ring + slot * kSlotBytes + c * kChunkBytes, img + ((size_t)c * n + h0) * 128u, bytes, bar_full + 8 * slot);
                    }
                    __syncwarp();
                    if (++slot == kSlots) { slot = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: tiles A (slot 0) and B (slot 1) of the pair alternate part by part and share every weight slot
        uint32_t slot = 0, phase = 0, a_phase = 0;
        int dbg_n = lane == 0 ? 0 : kDbgCap;
        const uint64_t bdesc0 = umma_desc_sw128(s_ring);
        for (int pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
            const int nt = min(tps, p.num_tiles - tps * pair);
            for (int l = 0; l < n_layers; ++l) {
                const int n = p.layer[l].n, kch = p.layer[l].kchunks;
                for (int h0 = 0; h0 < n; h0 += kPartN) {
                    const uint32_t idesc = umma_idesc_bf16(kTileM, min(kPartN, n - h0));
                    const uint64_t bdesc = bdesc0 + (uint64_t)(slot * (kSlotBytes >> 4));
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        if (t < nt) {
                            mbar_wait(bar_a + 8 * t, (a_phase >> t) & 1u); a_phase ^= (1u << t);   // D(t) free; A(t) of this layer in TMEM
                            if (t == 0) mbar_wait(bar_full + 8 * slot, phase);
                            tc_fence_after();
                            TACO_DBG(0, dbg_n, 0x20 | t);
                            if (elect_one_sync()) {
                                const uint32_t a_addr = tmem0 + (uint32_t)(t * kTmemSlot);
                                const uint32_t d_addr = a_addr + (uint32_t)kTmemD;
                                for (int c = 0; c < kch; ++c) {
                                    const uint64_t b_c = bdesc + (uint64_t)(c * (kChunkBytes >> 4));
                                    const uint32_t a_c = a_addr + (uint32_t)(c * (kKC / 2));
                                    // 16 bf16 along K = 8 packed TMEM columns of A = 32 bytes inside the B swizzle atom
                                    umma_bf16_ts(d_addr, a_c, b_c, idesc, (uint32_t)(c != 0));
                                    umma_bf16_ts(d_addr, a_c + 8u, b_c + 2u, idesc, 1u);
                                    umma_bf16_ts(d_addr, a_c + 16u, b_c + 4u, idesc, 1u);
                                    umma_bf16_ts(d_addr, a_c + 24u, b_c + 6u, idesc, 1u);
                                }
                                if (t == nt - 1) umma_commit(bar_empty + 8 * slot);   // slot free once the last user's MMAs have read it
                                umma_commit(bar_d + 8 * t);
                            }
                            __syncwarp();
                            TACO_DBG(0, dbg_n, 0x30 | t);
                        }
                    }
                    if (++slot == kSlots) { slot = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ===================== epilogue warps: thread <-> TMEM lane <-> env row of tile slot t; the two 64-column halves of a
        // part are handled by the warps `ch` = 0 / 1 of the same lane quadrant
        const int e = warp - 2;
        const int t = e >> 3, ch = (e >> 2) & 1, quad = warp & 3;    // a warp may only touch TMEM lanes 32*(warp%4)..+31
        const int r = (quad << 5) | lane;
        const uint32_t t_a = tmem0 + ((uint32_t)(quad << 5) << 16) + (uint32_t)(t * kTmemSlot);
        const uint32_t t_d = t_a + (uint32_t)kTmemD + (uint32_t)(ch * 64);
        uint32_t d_phase = 0;
        const int kc0 = p.layer[0].kchunks;
        int dbg_n = (ch == 0 && quad == 0 && lane == 0) ? 0 : kDbgCap;
        int tile = t < tps ? tps * (int)blockIdx.x + t : p.num_tiles;      // tile slot 1 idles with one tile per CTA
        const int tile_step = tps * (int)gridDim.x;
        // Bulk-staged observations (p.obs_bulk): a FULL tile's rows are one contiguous block of global memory (128 x in_dim floats), so
        // one elected thread of the tile has the bulk-copy engine fetch the block of the tile AFTER the one being staged; the
        // 128 x 26 row-per-thread scalar loads (one cache line per lane and instruction) leave the LSU.  Partial last tiles use loads.
        float* s_obs_t = s_obs + t * (kTileM * kObsBulkMaxIn);
        const uint32_t obs_bytes = (uint32_t)(kTileM * p.in_dim * 4);
        uint32_t o_phase = 0;
        const bool staged_first = p.obs_bulk && (long long)(tile + 1) * kTileM <= p.n_rows;
        auto fetch_obs = [&](int tl) {                               // all 256 threads of the tile call this; one issues the copy
            asm volatile("bar.sync %0, %1;" ::"r"(1 + t), "r"((kEpiWarps / 2) * 32) : "memory");     // every reader of the buffer is done
            if (e % 8 == 0 && elect_one_sync()) {
                fence_proxy_async_smem();
                mbar_arrive_expect_tx(bar_obs + 8 * t, obs_bytes);
                bulk_g2s(smem_u32(s_obs_t), p.obs + (size_t)tl * kTileM * p.in_dim, obs_bytes, bar_obs + 8 * t);
            }
        };
        auto bulk_tile = [&](int tl) { return p.obs_bulk && tl < p.num_tiles && (long long)(tl + 1) * kTileM <= p.n_rows; };
        if (tile < p.num_tiles) {
            // ---- the first tile's observation row as packed bf16 into A(t), K padded with zeros to kc0 * 64; groups of 32
            // features (16 TMEM columns) alternate between ch 0 / 1
            const long long row0 = (long long)tile * kTileM + r;
            if (staged_first) {
                fetch_obs(tile);
                mbar_wait(bar_obs + 8 * t, o_phase); o_phase ^= 1u;
                uint32_t pk[16];
                if (ch == 0) read_staged_obs(s_obs_t, r, p.in_dim, pk);
                else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = 0u;
                }
                tmem_st16(t_a + (uint32_t)(ch * 16), pk);
                if (bulk_tile(tile + tile_step)) fetch_obs(tile + tile_step);
            } else {
                for (int g = ch; g < kc0 * 2; g += 2) {
                    uint32_t pk[16];
                    load_obs_group(p.obs + row0 * p.in_dim, g, p.in_dim, row0 < p.n_rows, pk, p.obs_vec2 != 0);
                    tmem_st16(t_a + (uint32_t)(g * 16), pk);
                }
                if (bulk_tile(tile + tile_step)) fetch_obs(tile + tile_step);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_a + 8 * t);
            TACO_DBG(1 + t, dbg_n, 0x01);
        }
        for (; tile < p.num_tiles; tile += tile_step) {
            const long long row = (long long)tile * kTileM + r;
            const bool valid = row < p.n_rows;
            const bool has_next = tile + tile_step < p.num_tiles;
            const long long row_next = row + (long long)tile_step * kTileM;
            const float* x_next = p.obs + row_next * p.in_dim;
            if (has_next && row_next < p.n_rows) prefetch_l2(x_next + ch * 32);   // this thread's first group of the next tile
            for (int l = 0; l < p.n_hidden; ++l) {
                // hidden layer: relu(D + bias) as bf16 pairs, held in registers until every MMA of the layer has read A(t), then
                // written over A(t) as the next layer's operand
                const int n = p.layer[l].n;
                const float* bl = s_bias + l * kMaxN;
                const bool two_parts = n > kPartN;
                uint32_t pk0[32];
                const bool mine0 = ch * 64 < min(n, kPartN);                 // this warp has columns in part 0
                const bool mine1 = two_parts && (kPartN + ch * 64 < n);      // ... in part 1
                mbar_wait(bar_d + 8 * t, d_phase); d_phase ^= 1u;
                tc_fence_after();
                TACO_DBG(1 + t, dbg_n, 0x02);
                if (mine0) {
                    uint32_t v[32];
                    tmem_ld32(t_d, v); tmem_ld_wait();
                    relu_pack32(v, bl + ch * 64, pk0);
                    tmem_ld32(t_d + 32u, v); tmem_ld_wait();
                    if (two_parts) { tc_fence_before(); mbar_arrive(bar_a + 8 * t); TACO_DBG(1 + t, dbg_n, 0x03); }   // D drained: part 1 may start
                    relu_pack32(v, bl + ch * 64 + 32, pk0 + 16);
                } else if (two_parts) { tc_fence_before(); mbar_arrive(bar_a + 8 * t); }
                if (two_parts) {
                    mbar_wait(bar_d + 8 * t, d_phase); d_phase ^= 1u;
                    tc_fence_after();
                    TACO_DBG(1 + t, dbg_n, 0x04);
                }
                // every MMA of layer l on this tile is complete: overwrite A(t) (feature k -> packed column k/2)
                if (mine0) { tmem_st16(t_a + (uint32_t)(ch * 32), pk0); tmem_st16(t_a + (uint32_t)(ch * 32 + 16), pk0 + 16); }
                if (mine1) {
                    uint32_t v[32], pk1[16];
                    tmem_ld32(t_d, v); tmem_ld_wait();
                    relu_pack32(v, bl + kPartN + ch * 64, pk1);
                    tmem_st16(t_a + (uint32_t)(64 + ch * 32), pk1);
                    tmem_ld32(t_d + 32u, v); tmem_ld_wait();
                    relu_pack32(v, bl + kPartN + ch * 64 + 32, pk1);
                    tmem_st16(t_a + (uint32_t)(64 + ch * 32 + 16), pk1);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(bar_a + 8 * t);
                TACO_DBG(1 + t, dbg_n, 0x05);
            }
            // ---- output layer (columns 0..3 of its 16-wide part).  The next tile's first observation group is fetched while the
            // output MMAs run; once D is in registers the next A(t) is written and the issuer released, and only then the tanh /
            // sampling tail of this tile runs (overlapping the next tile's first MMAs).
            uint32_t pkn[16];
            const bool next_staged = has_next && bulk_tile(tile + tile_step);
            if (next_staged) {                                       // the block was fetched a whole tile ago: no load latency to hide
                mbar_wait(bar_obs + 8 * t, o_phase); o_phase ^= 1u;
                if (ch == 0) read_staged_obs(s_obs_t, r, p.in_dim, pkn);
                else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) pkn[j] = 0u;
                }
                if (bulk_tile(tile + 2 * tile_step)) fetch_obs(tile + 2 * tile_step);
            } else if (has_next) load_obs_group(x_next, ch, p.in_dim, row_next < p.n_rows, pkn, p.obs_vec2 != 0);
            mbar_wait(bar_d + 8 * t, d_phase); d_phase ^= 1u;
            tc_fence_after();
            TACO_DBG(1 + t, dbg_n, 0x06);
            uint32_t v[4];
            if (ch == 0) { tmem_ld4(t_d, v); tmem_ld_wait(); }
            if (has_next) {
                tmem_st16(t_a + (uint32_t)(ch * 16), pkn);
                for (int g = ch + 2; g < kc0 * 2 && !next_staged; g += 2) {
                    load_obs_group(x_next, g, p.in_dim, row_next < p.n_rows, pkn, p.obs_vec2 != 0);
                    tmem_st16(t_a + (uint32_t)(g * 16), pkn);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(bar_a + 8 * t);
                TACO_DBG(1 + t, dbg_n, 0x01);
            }
            if (ch == 0 && valid) {
                float pre[kOutPad];
#pragma unroll
                for (int o = 0; o < kOutPad; ++o) pre[o] = __uint_as_float(v[o]) + __ldg(p.b_out + o);
                actor_tail(pre, p.out_dim, row, p.mean, p.sp);
            }
            TACO_DBG(1 + t, dbg_n, 0x07);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem0, kTmemCols);
    }
}

#endif  // TACO_TC_NO_ACTOR_KERNEL

// fp32 (out, in) row-major weights -> bf16 K-chunk images in the SWIZZLE_128B K-major layout the kernel copies verbatim:
// chunk c holds W[:, 64c .. 64c+63] as n rows of 128 bytes (rows >= n_real are zero); the 16-byte piece q of row r sits
// at r*128 + ((q ^ (r&7)) << 4).
static __global__ void pack_weights_kernel(const float* __restrict__ w, int n_real, int n, int k, uint8_t* __restrict__ img) {
    const int kchunks = (k + kKC - 1) / kKC;
    const int total = kchunks * n * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int q = i & 7, r = (i >> 3) % n, c = (i >> 3) / n;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kk = c * kKC + q * 8 + j;
            f[j] = (kk < k && r < n_real) ? w[(size_t)r * k + kk] : 0.0f;
        }
        const uint4 pk = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        *reinterpret_cast<uint4*>(img + (size_t)c * n * 128 + sw128_off(r, q)) = pk;
    }
}

}  // namespace actor
}  // namespace taco

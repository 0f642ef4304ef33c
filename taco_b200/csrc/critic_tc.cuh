// Critic (LSTM encoder over the state history + MLP -> value) on the 5th-generation tensor cores, one persistent CTA per SM.
//
// Replaces, for rollout inference, the critic branch of PPO_ActorCritic.act
// (IsaacGymEnvs/algorithms/nets_asymmetry.py:350-352): value = critic_mlp(critic_encoder(states)) with
//   LSTMEncoder.forward   nets_asymmetry.py:128-136   nn.LSTM(26, H, 1, batch_first=True) over (N, len_states, 26), output h_T
//   MLP.forward           nets_asymmetry.py:23-39     [Linear -> ReLU] x L -> Linear(-> 1), Identity output (:318)
//
// Same machinery as the actor kernel (actor_tc.cuh: two 128-env tiles in flight per CTA, the A operand in TENSOR MEMORY,
// weights streamed through a shared-memory ring by the bulk-copy engine and shared by both tiles, MMA parts of the two
// tiles interleaved so that one tile's accumulator is drained while the tensor core works on the other).  What is new is
// the schedule and the LSTM epilogue:
//   * one LSTM time step is ONE layer of the chain: gates[128 x 4H] = [h_{t-1} | x_t] (K = 64 + 64, zero padded) * Wcat^T,
//     Wcat = [W_hh | W_ih] with its rows permuted so that every 16 accumulator columns hold the four gates (i, f, g, o; 4
//     columns each) of the same 4 hidden units -- an epilogue thread reads 16 columns and owns those 4 units outright;
//   * the cell state c never leaves the registers of the thread that owns (env row, hidden unit): 2 * H / 4 values per
//     thread; h_t is written back over the A operand as packed bf16 pairs next to the freshly staged x_{t+1};
//   * the T = len_states steps reuse one weight image; then the MLP layers and the 16-wide output part follow as in the actor.
// Gate non-linearities use MUFU.TANH (tanh.approx; sigmoid(x) = 0.5 + 0.5 tanh(x / 2)): 5 MUFU per (env, unit, step).
#pragma once
#define TACO_TC_NO_ACTOR_KERNEL
#include "actor_tc.cuh"

namespace taco {
namespace critic {
using namespace taco::actor;

constexpr int kMaxSeq = 8;             // len_states of the reference runs is 5 (README.md:60-66)
constexpr int kTcMaxLstmHidden = 64;   // c lives in registers: H / 2 floats per epilogue thread
constexpr int kTcMaxIn = 32;           // features per frame (26) padded to 32 = 16 packed TMEM columns
constexpr int kMaxMlpHidden = 3;       // s_bias rows: LSTM gates + up to 3 hidden layers
constexpr int kAColH = 0;              // A operand columns (packed bf16 pairs): h in [0, 32), x_t in [32, 48), zero pad [48, 64)
constexpr int kAColX = 32;

struct CriticTcParams {
    const float* states;   // (n_rows, seq_len, in_dim) f32
    float* value;          // (n_rows) f32
    int in_dim, seq_len, n_rows, num_tiles;
    int lstm_hidden;       // H: multiple of 16, <= 64
    int n_hidden;          // MLP hidden layers, 1..3
    const uint8_t* wimg;   // pre-swizzled bf16 chunk images: LSTM [W_hh | W_ih] (gate-permuted rows), MLP layers, output (16 rows)
    const float* bias;     // [kMaxHidden][kMaxN]: row 0 = b_ih + b_hh in the permuted column order, rows 1.. = MLP hidden biases
    const float* b_out;    // [kOutPad]
    TcLayer layer[2 + kMaxMlpHidden];   // [0] one LSTM step, [1 .. n_hidden] MLP hidden layers, [n_hidden + 1] output layer
};

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
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
__device__ __forceinline__ float tanh_mufu(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_mufu(float x) { return fmaf(0.5f, tanh_mufu(0.5f * x), 0.5f); }

// 16 accumulator columns = [i(4) f(4) g(4) o(4)] of 4 hidden units: c' = sig(f) c + sig(i) tanh(g), h' = sig(o) tanh(c');
// h' leaves as 2 packed bf16 pairs (unit 2j in the low half)
__device__ __forceinline__ void lstm_group(const uint32_t (&v)[16], const float* bias, float* c, uint32_t* hp) {
#pragma unroll
    for (int u = 0; u < 4; u += 2) {
        float h2[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int k = u + e;
            const float ig = sigmoid_mufu(__uint_as_float(v[k]) + bias[k]);
            const float fg = sigmoid_mufu(__uint_as_float(v[4 + k]) + bias[4 + k]);
            const float gg = tanh_mufu(__uint_as_float(v[8 + k]) + bias[8 + k]);
            const float og = sigmoid_mufu(__uint_as_float(v[12 + k]) + bias[12 + k]);
            c[k] = fmaf(fg, c[k], ig * gg);
            h2[e] = og * tanh_mufu(c[k]);
        }
        hp[u >> 1] = pack_bf16x2(h2[0], h2[1]);
    }
}
// one 64-column half of a gate part = 4 groups = 16 hidden units of this thread's env row; `early` = arrive on the tile's
// A/D barrier as soon as the last accumulator columns are in registers (the next part of the layer may then overwrite D)
__device__ __forceinline__ void lstm_half(uint32_t t_d, const float* bias, float* c, uint32_t* hp, bool early, uint32_t bar) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t v[16];
        tmem_ld16(t_d + (uint32_t)(g * 16), v); tmem_ld_wait();
        if (g == 3 && early) { tc_fence_before(); mbar_arrive(bar); }
        lstm_group(v, bias + g * 16, c + g * 4, hp + g * 2);
    }
}

// features [16 ch, 16 ch + 16) of one state frame as 8 packed bf16 pairs (zero beyond in_dim / for invalid rows)
__device__ __forceinline__ void load_x16(const float* x, int ch, int in_dim, bool valid, uint32_t* pk) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = ch * 16 + 2 * j;
        const float f0 = (valid && k < in_dim) ? __ldg(x + k) : 0.0f;
        const float f1 = (valid && k + 1 < in_dim) ? __ldg(x + k + 1) : 0.0f;
        pk[j] = pack_bf16x2(f0, f1);
    }
}

// A operand of a fresh tile: h_0 = 0 in columns [0, 32), x_0 in [32, 48), zeros in [48, 64) (TMEM is never read uninitialised:
// the weight rows of the padding are zero, but 0 * NaN is NaN)
__device__ __forceinline__ void stage_new_tile(uint32_t t_a, int ch, const uint32_t* x0) {
    uint32_t z[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) z[j] = 0u;
    tmem_st16(t_a + (uint32_t)(kAColH + ch * 16), z);
    tmem_st8(t_a + (uint32_t)(kAColX + ch * 8), x0);
    tmem_st8(t_a + (uint32_t)(kAColX + 16 + ch * 8), z);
}

__global__ void __launch_bounds__(kTcThreads, 1) critic_tc_kernel(const CriticTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const uint32_t s_ring = base;
    float* s_bias = reinterpret_cast<float*>(sm + kSlots * kSlotBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + kMaxHidden * kMaxN);
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = bar_full + 8 * kSlots;
    const uint32_t bar_a = bar_empty + 8 * kSlots;
    const uint32_t bar_d = bar_a + 16;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kSlots + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_pairs = (p.num_tiles + 1) >> 1;
    const int T = p.seq_len;
    const int n_sched = T + p.n_hidden + 1;            // T LSTM steps, the MLP hidden layers, the output part

    for (int i = threadIdx.x; i < kMaxHidden * kMaxN; i += kTcThreads) s_bias[i] = p.bias[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlots; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(bar_a + 8 * t, (kEpiWarps / 2) * 32); mbar_init(bar_d + 8 * t, 1); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(s_tmem), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *s_tmem;

    if (warp == 0) {
        // ===================== weight producer: per pair, per schedule step, per 128-row part: one ring slot
        uint32_t slot = 0, phase = 0;
        for (int pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
            for (int s = 0; s < n_sched; ++s) {
                const int l = s < T ? 0 : s - T + 1;
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
        // ===================== MMA issuer (see actor_tc.cuh): the two tiles alternate part by part and share every weight slot
        uint32_t slot = 0, phase = 0, a_phase = 0;
        const uint64_t bdesc0 = umma_desc_sw128(s_ring);
        for (int pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
            const int nt = (2 * pair + 1 < p.num_tiles) ? 2 : 1;
            for (int s = 0; s < n_sched; ++s) {
                const int l = s < T ? 0 : s - T + 1;
                const int n = p.layer[l].n, kch = p.layer[l].kchunks;
                for (int h0 = 0; h0 < n; h0 += kPartN) {
                    const uint32_t idesc = umma_idesc_bf16(kTileM, min(kPartN, n - h0));
                    const uint64_t bdesc = bdesc0 + (uint64_t)(slot * (kSlotBytes >> 4));
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        if (t < nt) {
                            mbar_wait(bar_a + 8 * t, (a_phase >> t) & 1u); a_phase ^= (1u << t);
                            if (t == 0) mbar_wait(bar_full + 8 * slot, phase);
                            tc_fence_after();
                            if (elect_one_sync()) {
                                const uint32_t a_addr = tmem0 + (uint32_t)(t * kTmemSlot);
                                const uint32_t d_addr = a_addr + (uint32_t)kTmemD;
                                for (int c = 0; c < kch; ++c) {
                                    const uint64_t b_c = bdesc + (uint64_t)(c * (kChunkBytes >> 4));
                                    const uint32_t a_c = a_addr + (uint32_t)(c * (kKC / 2));
                                    umma_bf16_ts(d_addr, a_c, b_c, idesc, (uint32_t)(c != 0));
                                    umma_bf16_ts(d_addr, a_c + 8u, b_c + 2u, idesc, 1u);
                                    umma_bf16_ts(d_addr, a_c + 16u, b_c + 4u, idesc, 1u);
                                    umma_bf16_ts(d_addr, a_c + 24u, b_c + 6u, idesc, 1u);
                                }
                                if (t == nt - 1) umma_commit(bar_empty + 8 * slot);
                                umma_commit(bar_d + 8 * t);
                            }
                            __syncwarp();
                        }
                    }
                    if (++slot == kSlots) { slot = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ===================== epilogue warps: thread <-> TMEM lane <-> env row of tile slot t; `ch` = 64-column half of a part
        const int e = warp - 2;
        const int t = e >> 3, ch = (e >> 2) & 1, quad = warp & 3;
        const int r = (quad << 5) | lane;
        const uint32_t t_a = tmem0 + ((uint32_t)(quad << 5) << 16) + (uint32_t)(t * kTmemSlot);
        const uint32_t t_d = t_a + (uint32_t)kTmemD + (uint32_t)(ch * 64);
        uint32_t d_phase = 0;
        const size_t row_floats = (size_t)T * p.in_dim;
        int tile = 2 * (int)blockIdx.x + t;
        const int tile_step = 2 * (int)gridDim.x;
        if (tile < p.num_tiles) {
            const long long row0 = (long long)tile * kTileM + r;
            uint32_t x0[8];
            load_x16(p.states + row0 * row_floats, ch, p.in_dim, row0 < p.n_rows, x0);
            stage_new_tile(t_a, ch, x0);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(bar_a + 8 * t);
        }
        // the LSTM layer: n = 4H accumulator columns in parts of 128; this thread owns 16 hidden units of every part it has columns in
        const int n_g = p.layer[0].n;
        const bool g_two = n_g > kPartN;
        const bool g_mine0 = ch * 64 < min(n_g, kPartN);
        const bool g_mine1 = g_two && (kPartN + ch * 64 < n_g);
        for (; tile < p.num_tiles; tile += tile_step) {
            const long long row = (long long)tile * kTileM + r;
            const bool valid = row < p.n_rows;
            const bool has_next = tile + tile_step < p.num_tiles;
            const long long row_next = row + (long long)tile_step * kTileM;
            const float* xs = p.states + row * row_floats;
            const float* xs_next = p.states + row_next * row_floats;
            if (has_next && row_next < p.n_rows) prefetch_l2(xs_next + ch * 16);
            float c[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) c[j] = 0.0f;
            // ---- T LSTM steps
            for (int ts = 0; ts < T; ++ts) {
                const bool have_x = ts + 1 < T;
                uint32_t xn[8], hp[16];
                if (have_x) load_x16(xs + (size_t)(ts + 1) * p.in_dim, ch, p.in_dim, valid, xn);
                mbar_wait(bar_d + 8 * t, d_phase); d_phase ^= 1u;
                tc_fence_after();
                if (g_mine0) lstm_half(t_d, s_bias + ch * 64, c, hp, g_two, bar_a + 8 * t);      // D drained: part 1 may start
                else if (g_two) { tc_fence_before(); mbar_arrive(bar_a + 8 * t); }
                if (g_two) {
                    mbar_wait(bar_d + 8 * t, d_phase); d_phase ^= 1u;
                    tc_fence_after();
                }
                if (g_mine1) lstm_half(t_d, s_bias + kPartN + ch * 64, c + 16, hp + 8, false, 0u);
                // every MMA of this step on this tile is complete: h_t over h_{t-1} (unit k -> packed column k / 2), x_{t+1} over x_t
                if (g_mine0) tmem_st8(t_a + (uint32_t)(kAColH + ch * 8), hp);
                if (g_mine1) tmem_st8(t_a + (uint32_t)(kAColH + 16 + ch * 8), hp + 8);
                if (have_x) tmem_st8(t_a + (uint32_t)(kAColX + ch * 8), xn);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(bar_a + 8 * t);
            }
            // ---- MLP hidden layers (as in the actor kernel)
            for (int l = 1; l <= p.n_hidden; ++l) {
                const int n = p.layer[l].n;
                const float* bl = s_bias + l * kMaxN;
                const bool two_parts = n > kPartN;
                uint32_t pk0[32];
                const bool mine0 = ch * 64 < min(n, kPartN);
                const bool mine1 = two_parts && (kPartN + ch * 64 < n);
                mbar_wait(bar_d + 8 * t, d_phase); d_phase ^= 1u;
                tc_fence_after();
                if (mine0) {
                    uint32_t v[32];
                    tmem_ld32(t_d, v); tmem_ld_wait();
                    relu_pack32(v, bl + ch * 64, pk0);
                    tmem_ld32(t_d + 32u, v); tmem_ld_wait();
                    if (two_parts) { tc_fence_before(); mbar_arrive(bar_a + 8 * t); }
                    relu_pack32(v, bl + ch * 64 + 32, pk0 + 16);
                } else if (two_parts) { tc_fence_before(); mbar_arrive(bar_a + 8 * t); }
                if (two_parts) {
                    mbar_wait(bar_d + 8 * t, d_phase); d_phase ^= 1u;
                    tc_fence_after();
                }
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
            }
            // ---- output part (column 0 of 16): the next tile's A operand is staged before the value leaves
            uint32_t x0[8];
            if (has_next) load_x16(xs_next, ch, p.in_dim, row_next < p.n_rows, x0);
            mbar_wait(bar_d + 8 * t, d_phase); d_phase ^= 1u;
            tc_fence_after();
            uint32_t v[4];
            if (ch == 0) { tmem_ld4(t_d, v); tmem_ld_wait(); }
            if (has_next) {
                stage_new_tile(t_a, ch, x0);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(bar_a + 8 * t);
            }
            if (ch == 0 && valid) p.value[row] = __uint_as_float(v[0]) + __ldg(p.b_out);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem0, kTmemCols);
    }
}

// LSTM weights (torch layout: w_ih (4H, In), w_hh (4H, H), gate row blocks i, f, g, o) -> two bf16 K-chunk images of n = 4H
// rows each in the SWIZZLE_128B K-major layout: chunk 0 = W_hh (K padded to 64), chunk 1 = W_ih (K padded to 64).  Image row
// r = 128 part + j holds gate (j % 16) / 4 of hidden unit 32 part + 4 (j / 16) + j % 4.  bias_perm[r] = b_ih + b_hh of that row.
__device__ __forceinline__ int lstm_src_row(int r, int hidden) {
    const int part = r >> 7, j = r & 127;
    const int unit = part * 32 + (j >> 4) * 4 + (j & 3), gate = (j & 15) >> 2;
    return unit < hidden ? gate * hidden + unit : -1;
}
__global__ void pack_lstm_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                                 const float* __restrict__ b_hh, int hidden, int in_dim, int n, uint8_t* __restrict__ img,
                                 float* __restrict__ bias_perm) {
    const int total = 2 * n * 8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int q = i & 7, r = (i >> 3) % n, c = (i >> 3) / n;
        const int src = lstm_src_row(r, hidden);
        const float* w = c == 0 ? w_hh : w_ih;
        const int k = c == 0 ? hidden : in_dim;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kk = q * 8 + j;
            f[j] = (src >= 0 && kk < k) ? w[(size_t)src * k + kk] : 0.0f;
        }
        const uint4 pk = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
        *reinterpret_cast<uint4*>(img + (size_t)c * n * 128 + sw128_off(r, q)) = pk;
        if (c == 0 && q == 0) bias_perm[r] = src >= 0 ? b_ih[src] + b_hh[src] : 0.0f;
    }
}

}  // namespace critic
}  // namespace taco

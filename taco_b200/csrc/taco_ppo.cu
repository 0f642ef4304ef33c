// Native PPO update behind the C ABI (include/taco_b200.h, taco_ppo_*): SURVEY.md section 8f row 4.
//
// Mirrors PPO.update of the reference (IsaacGymEnvs/algorithms/ppo_asymmetry.py:137-258) for the network configuration the
// reference trains with (README.md:60-66; nets_asymmetry.py:270-377): actor = MLP(obs) -> tanh mean, state-independent log_std,
// MultivariateNormal(mean, scale_tril = diag(exp(log_std)^2)); critic = MLP(LSTMEncoder(states)).  Per minibatch:
//
//   gather        obs / states / action / old log-prob / advantage / return rows of the minibatch indices -> bf16 operands
//   forward       actor MLP and critic (5 LSTM steps + MLP) as tcgen05 GEMMs (gemm_tc.cuh), activations saved for the backward pass
//   loss          PPO_ActorCritic.evaluate + the loss terms of :190-212: ratio, clipped surrogate, value MSE, entropy, approx KL
//                 (:218-221); writes d loss / d mean-pre-activation and d loss / d value, accumulates the statistics
//   decide        the KL early stop of :223-226 taken ON THE DEVICE (a flag later launches test), so an update needs no host sync
//   backward      dX / dW GEMMs (split-K partials for the K = batch weight gradients), LSTM backward through time
//   apply         gradient assembly + global norm (clip_grad_norm_, :244), Adam (torch.optim.Adam, eps 1e-5, :117), spectral
//                 projection of the actor weights (:248-249, :398-404), re-pack of the bf16 operand copies
//
// Arithmetic: bf16 operands, fp32 accumulation (tensor memory), fp32 master weights / Adam moments / losses.  Everything
// else on the path (gathers, losses, LSTM point-wise backward, reductions, Adam) is plain fp32 CUDA.
#include "launch_count.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/taco_b200.h"
#include "gemm_tc.cuh"

namespace taco {
int fail(int code, const std::string& msg);      // taco_env.cu
namespace ppo {

using namespace taco::gemm;
typedef __nv_bfloat16 bf16;

#define PPO_CUDA(expr)                                                                               \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) return ::taco::fail(TACO_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

constexpr int kMaxHiddenL = 4;
constexpr int kH = 64;              // LSTM hidden width the LSTM epilogue is written for
constexpr int kG = 4 * kH;          // gate rows
constexpr int kSeqMax = 8;
constexpr int kNumStat = 16;        // accumulators: [0] surrogate, [1] value, [2] kl, [3] count, [4..7] d log_std, [8] grad norm^2
constexpr int kLogCols = 8;         // per minibatch log row: pg, value, entropy, total, kl, grad_norm, stopped, clip_coef

static inline int pad8(int x) { return (x + 7) & ~7; }
static inline int pad16(int x) { return (x + 15) & ~15; }

// ---------------------------------------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}
// row-major bf16 matrix [rows][cols], row pitch ld elements; box = 64 columns (one 128-byte swizzle row) x box_rows rows
static bool make_map(CUtensorMap* tm, const bf16* ptr, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn f = encode_fn();
    if (!f) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(bf16)};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return f(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp32 [rows][cols] output of the split-K GEMMs, stored tile by tile (32 columns x 128 rows) by the TMA store engine
static bool make_map_f32(CUtensorMap* tm, const float* ptr, long long rows, long long cols, long long ld) {
    EncodeTiledFn f = encode_fn();
    if (!f) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {32u, (cuuint32_t)BM};
    const cuuint32_t estr[2] = {1, 1};
    return f(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---------------------------------------------------------------------------------------------- small kernels
struct Hyper {
    float lr, clip, target_kl, max_grad, pi_coef, vf_coef, ent_coef, lipschitz;
    int use_lipschitz, world;
};

// minibatch gather: rows idx[i] of the rollout tensors -> bf16 operands (batch-major) + compact fp32 side arrays
struct GatherParams {
    const float* obs; int obs_dim;                  // [N][obs_dim]
    const float* states; int seq, state_dim;        // [N][seq][state_dim]
    const float* act; const float* logp; const float* adv; const float* ret;   // [N][A], [N], [N], [N]
    const long long* idx; int B, act_dim;
    bf16* x0_bm; int ld_x0;                         // [B][ld_x0]
    bf16* u_bm; int ld_u;                           // [seq*B][ld_u]: columns [h(64) | x | 1 1 | 0]
    float* g_act; float* g_logp; float* g_adv; float* g_ret;
    const int* stop;
};
// one operand matrix for the 32 samples of a CTA: features [0, n_src) come from the sample's source row, [n_src, n_src + n_one) are
// the constant 1, the rest up to n_out is 0
__device__ __forceinline__ void gather_matrix(const float* src, long long src_stride, const long long* rows, int n_valid,
                                              int n_src, int n_one, int n_out, bf16* bm, long long ld_bm) {
    for (int e = threadIdx.x; e < n_valid * n_out; e += blockDim.x) {
        const int sidx = e / n_out, f = e % n_out;
        const float v = f < n_src ? __ldg(src + rows[sidx] * src_stride + f) : (f < n_src + n_one ? 1.0f : 0.0f);
        bm[(long long)sidx * ld_bm + f] = __float2bfloat16_rn(v);
    }
}
__global__ void __launch_bounds__(256) gather_kernel(const GatherParams p) {
    pdl_launch_dependents();
    pdl_wait();
    if (p.stop && *p.stop) return;
    __shared__ long long rows[32];
    const int i0 = blockIdx.x * 32;
    const int nv = min(32, p.B - i0);
    if (threadIdx.x < 32) rows[threadIdx.x] = threadIdx.x < nv ? p.idx[i0 + threadIdx.x] : 0;
    __syncthreads();
    gather_matrix(p.obs, p.obs_dim, rows, nv, p.obs_dim, 0, (p.obs_dim + 7) & ~7, p.x0_bm + (long long)i0 * p.ld_x0, p.ld_x0);
    // U_t = [h_{t-1} (64, written by the previous LSTM step; h_0 = 0 stays as allocated) | x_t | 1 1 | 0]: two constant-1 inputs carry the bias
    for (int t = 0; t < p.seq; ++t)
        gather_matrix(p.states + (long long)t * p.state_dim, (long long)p.seq * p.state_dim, rows, nv, p.state_dim, 2, p.ld_u - kH,
                      p.u_bm + ((long long)t * p.B + i0) * p.ld_u + kH, p.ld_u);
    for (int e = threadIdx.x; e < nv * (p.act_dim + 3); e += blockDim.x) {
        const int sidx = e / (p.act_dim + 3), c = e % (p.act_dim + 3);
        const long long r = rows[sidx];
        if (c < p.act_dim) p.g_act[(long long)(i0 + sidx) * p.act_dim + c] = __ldg(p.act + r * p.act_dim + c);
        else if (c == p.act_dim) p.g_logp[i0 + sidx] = __ldg(p.logp + r);
        else if (c == p.act_dim + 1) p.g_adv[i0 + sidx] = __ldg(p.adv + r);
        else p.g_ret[i0 + sidx] = __ldg(p.ret + r);
    }
}

// PPO_ActorCritic.evaluate + losses (nets_asymmetry.py:356-377, ppo_asymmetry.py:190-221), one thread per sample
struct LossParams {
    const float* mean; const float* value;          // [B][A], [B]
    const float* act; const float* old_logp; const float* adv; const float* ret;
    const float* log_std;                           // [A] (master parameter)
    int B, act_dim;
    Hyper h;
    bf16* dz_bm;                                    // actor output-layer pre-activation gradient: [B][16]
    bf16* dv_bm;                                    // critic value gradient: [B][16] (column 0)
    double* acc;                                    // kNumStat accumulators
    const int* stop;
};
__global__ void __launch_bounds__(256) loss_kernel(const LossParams p) {
    pdl_launch_dependents();
    pdl_wait();
    if (p.stop && *p.stop) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float s_sur = 0.f, s_val = 0.f, s_kl = 0.f, dls[4] = {0.f, 0.f, 0.f, 0.f};
    if (i < p.B) {
        const int A = p.act_dim;
        float z[4], mu[4], inv_sd[4];
        float logp = -0.5f * A * 1.8378770664093453f;                  // -k/2 log(2 pi)
        for (int k = 0; k < A; ++k) {
            const float lsd = 2.0f * __ldg(p.log_std + k);              // std = exp(log_std)^2 (nets_asymmetry.py:338)
            inv_sd[k] = __expf(-lsd);
            mu[k] = p.mean[(long long)i * A + k];
            z[k] = (p.act[(long long)i * A + k] - mu[k]) * inv_sd[k];
            logp += -0.5f * z[k] * z[k] - lsd;
        }
        const float adv = p.adv[i];
        const float log_ratio = logp - p.old_logp[i];
        const float ratio = __expf(log_ratio);
        const float rc = fminf(fmaxf(ratio, 1.0f - p.h.clip), 1.0f + p.h.clip);
        const float s1 = adv * ratio, s2 = adv * rc;
        s_sur = -fminf(s1, s2);                                         // :196-199
        // d(-min(s1, s2)) / d logp: the unclipped branch when it is the minimum (ties: clamp passes the gradient inside the range)
        const float dl_dlogp = (s1 <= s2) ? -adv * ratio : 0.0f;
        const float dv = p.value[i] - p.ret[i];
        s_val = dv * dv;                                                // F.mse_loss(ret, value), :206
        s_kl = (ratio - 1.0f) - log_ratio;                              // :219-220
        const float invB = 1.0f / (float)p.B;
        const float gl = p.h.pi_coef * dl_dlogp * invB;
        for (int k = 0; k < 16; ++k) {
            float g = 0.0f;
            if (k < A) {
                g = gl * z[k] * inv_sd[k] * (1.0f - mu[k] * mu[k]);     // through mean = tanh(pre-activation)
                dls[k] = gl * 2.0f * (z[k] * z[k] - 1.0f);              // d logp / d log_std = 2 (z^2 - 1)
            }
            const bf16 b = __float2bfloat16_rn(g);
            p.dz_bm[(long long)i * 16 + k] = b;
            const bf16 bv = __float2bfloat16_rn(k == 0 ? p.h.vf_coef * 2.0f * dv * invB : 0.0f);
            p.dv_bm[(long long)i * 16 + k] = bv;
        }
    }
    __shared__ float red[7][8];
    float vals[7] = {s_sur, s_val, s_kl, dls[0], dls[1], dls[2], dls[3]};
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        float v = vals[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += (double)red[threadIdx.x][w];
        const int slot = threadIdx.x < 3 ? threadIdx.x : threadIdx.x + 1;       // 0,1,2 = losses; 4..7 = d log_std
        atomicAdd(p.acc + slot, s);
    }
}

// the per-minibatch bookkeeping of :214-226 on the device: means, log row, KL early stop
struct DecideParams {
    double* acc; float* log; int* n_logged; int* stop; int max_log;
    const float* log_std; int act_dim, B;
    Hyper h;
    float* grad_log_std;                           // flat gradient slot of log_std
};
__global__ void decide_kernel(const DecideParams p) {
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (*p.stop) return;
    const double w = (double)p.h.world, n = (double)p.B * w;           // under data parallelism acc holds the all-reduced sums
    const double pg = p.acc[0] / n, vl = p.acc[1] / n, kl = p.acc[2] / n;
    double ent = 0.5 * p.act_dim * (1.0 + 1.8378770664093453);
    for (int k = 0; k < p.act_dim; ++k) ent += 2.0 * (double)p.log_std[k];
    const double el = -ent;                                             // entropy_loss = -mean(entropy), :209
    const double total = p.h.pi_coef * pg + p.h.vf_coef * vl + p.h.ent_coef * el;
    const int row = *p.n_logged;
    if (row < p.max_log) {
        float* L = p.log + (size_t)row * kLogCols;
        L[0] = (float)pg; L[1] = (float)vl; L[2] = (float)el; L[3] = (float)total; L[4] = (float)kl; L[5] = 0.f; L[6] = 0.f; L[7] = 1.f;
    }
    *p.n_logged = row + 1;
    if (kl > 1.5 * (double)p.h.target_kl && p.h.pi_coef > 0.0f) {      // :223-226
        *p.stop = 1;
        if (row < p.max_log) p.log[(size_t)row * kLogCols + 6] = 1.f;
    }
    // d loss / d log_std (local sums: the flat gradient is averaged over ranks later like every other gradient)
    for (int k = 0; k < p.act_dim; ++k) p.grad_log_std[k] = (float)(p.acc[4 + k] / w) + p.h.ent_coef * -2.0f;
    for (int k = 0; k < 8; ++k) p.acc[k] = 0.0;
}

// LSTM backward, point-wise part of one time step (the gate math of nn.LSTM): one thread per (sample, unit), lanes along the units
struct LstmBwdParams {
    const float* dh; float* dc;                     // [B][64]: dL/dh_t (total), dL/dc_t carried from step t+1 (in) -> dL/dc_{t-1} (out)
    const bf16* gates; const float* c_prev; const float* c_cur;     // [B][256] activated gates of step t (chunked layout), c_{t-1} (null: 0), c_t
    bf16* dg_bm;                                    // [B][256] gate pre-activation gradients of step t
    int B; int first;                               // first = this is the last time step (dc_in = 0)
    const int* stop;
};
__global__ void __launch_bounds__(256) lstm_bwd_kernel(const LstmBwdParams p) {
    pdl_launch_dependents();
    pdl_wait();
    if (p.stop && *p.stop) return;
    const int j = threadIdx.x & (kH - 1);
    const long long i = (long long)blockIdx.x * 4 + (threadIdx.x >> 6);
    if (i >= p.B) return;
    const bf16* g = p.gates + i * kG + (j >> 4) * 64 + (j & 15);     // [4 chunks][4 gates][16 units] (gemm_tc.cuh LstmEpi::gates_out)
    const float gi = __bfloat162float(g[0]), gf = __bfloat162float(g[16]), gg = __bfloat162float(g[32]), go = __bfloat162float(g[48]);
    const float c = p.c_cur[i * kH + j], cp = p.c_prev ? p.c_prev[i * kH + j] : 0.0f;
    const float tc = tanhf(c);
    const float dh = p.dh[i * kH + j];
    const float dc = (p.first ? 0.0f : p.dc[i * kH + j]) + dh * go * (1.0f - tc * tc);
    p.dc[i * kH + j] = dc * gf;
    bf16* d = p.dg_bm + i * kG + j;
    d[0] = __float2bfloat16_rn(dc * gg * gi * (1.0f - gi));
    d[kH] = __float2bfloat16_rn(dc * cp * gf * (1.0f - gf));
    d[2 * kH] = __float2bfloat16_rn(dc * gi * (1.0f - gg * gg));
    d[3 * kH] = __float2bfloat16_rn(dh * tc * go * (1.0f - go));
}

// column sums of batch-major bf16 matrices = bias gradients.  A CTA sums kColRows rows of one matrix (lanes along the columns, bf16
// pairs) and writes one partial row; grad_assemble_kernel adds the partial rows in a fixed order (bit-reproducible).
constexpr int kColRows = 512, kColLd = 256;
struct ColSumSeg { const bf16* src; int ld; };      // [B][ld], ld a multiple of 16, <= 256
struct ColSumParams { ColSumSeg seg[2 * (kMaxHiddenL + 1)]; int n_seg; int B; float* partial; const int* stop; };   // partial: [seg][chunk][kColLd]
__global__ void __launch_bounds__(256) colsum_kernel(const ColSumParams p) {
    pdl_launch_dependents();
    pdl_wait();
    if (p.stop && *p.stop) return;
    __shared__ float2 part[256];
    const ColSumSeg sg = p.seg[blockIdx.y];
    const int pairs = sg.ld >> 1;                    // 8 .. 128 column pairs
    const int lanes = 256 / pairs;                   // row lanes
    const int cp = threadIdx.x % pairs, rl = threadIdx.x / pairs;
    const int r0 = blockIdx.x * kColRows, r1 = min(r0 + kColRows, p.B);
    float2 acc = make_float2(0.f, 0.f);
    if (rl < lanes)
        for (int r = r0 + rl; r < r1; r += lanes) {
            const uint32_t u = *reinterpret_cast<const uint32_t*>(sg.src + (long long)r * sg.ld + 2 * cp);
            acc.x += __uint_as_float(u << 16); acc.y += __uint_as_float(u & 0xFFFF0000u);
        }
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < pairs) {
        float2 t = make_float2(0.f, 0.f);
        for (int l = 0; l < lanes; ++l) { t.x += part[l * pairs + threadIdx.x].x; t.y += part[l * pairs + threadIdx.x].y; }
        float* dst = p.partial + ((long long)blockIdx.y * gridDim.x + blockIdx.x) * kColLd + 2 * threadIdx.x;
        dst[0] = t.x; dst[1] = t.y;
    }
}

// gradient assembly: flat_grad[dst + r * cols + c] = sum over splits of partial[s][r][col0 + c]; accumulates the global norm^2
struct GradSeg { const float* partial; int splits; long long split_stride; int ld, col0, rows, cols; long long dst; };
struct GradParams { GradSeg seg[32]; int n_seg; float* grad; double* acc; const int* stop; };
__global__ void __launch_bounds__(256) grad_assemble_kernel(const GradParams p) {
    pdl_launch_dependents();
    pdl_wait();
    if (p.stop && *p.stop) return;
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int s = 0; s < p.n_seg; ++s) {
        const long long cnt = (long long)p.seg[s].rows * p.seg[s].cols;
        if (e < cnt) {
            const GradSeg& g = p.seg[s];
            const int r = (int)(e / g.cols), c = (int)(e % g.cols);
            const float* src = g.partial + (long long)r * g.ld + g.col0 + c;
            float acc = 0.f;
            for (int k = 0; k < g.splits; ++k) acc += src[(long long)k * g.split_stride];
            p.grad[g.dst + e] = acc;
            return;
        }
        e -= cnt;
    }
}
__global__ void __launch_bounds__(256) grad_norm_kernel(const float* grad, long long n, float inv_world, double* acc, const int* stop) {
    pdl_launch_dependents();
    pdl_wait();
    if (stop && *stop) return;
    double s = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const double g = (double)grad[e] * inv_world;
        s += g * g;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        atomicAdd(acc + 8, t);
    }
}

// clip_grad_norm_ (:244) + torch.optim.Adam (betas 0.9 / 0.999, eps 1e-5, :117): element-wise over the flat parameter vector
__global__ void __launch_bounds__(256) adam_kernel(float* prm, float* m, float* v, const float* grad, long long n, Hyper h, int* step, double* acc,
                                                   float* log, const int* n_logged, int max_log, const int* stop) {
    pdl_launch_dependents();
    pdl_wait();
    if (stop && *stop) return;
    const double norm = sqrt(acc[8]);
    const float coef = fminf((float)((double)h.max_grad / (norm + 1e-6)), 1.0f) / (float)h.world;     // the 1 / world averages the summed gradient
    const int t = *step + 1;
    const float b1 = 0.9f, b2 = 0.999f;
    const float bc1 = 1.0f - powf(b1, (float)t), bc2 = 1.0f - powf(b2, (float)t);
    const float step_size = h.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const float g = grad[e] * coef;
        const float mm = b1 * m[e] + (1.0f - b1) * g;
        const float vv = b2 * v[e] + (1.0f - b2) * g * g;
        m[e] = mm; v[e] = vv;
        prm[e] -= step_size * (mm / (sqrtf(vv) * inv_sqrt_bc2 + 1e-5f));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const int row = *n_logged - 1;
        if (row >= 0 && row < max_log) { log[(size_t)row * kLogCols + 5] = (float)norm; log[(size_t)row * kLogCols + 7] = coef * (float)h.world; }
    }
}
__global__ void finish_step_kernel(int* step, double* acc, const int* stop) {
    pdl_launch_dependents();
    pdl_wait();
    if (stop && *stop) return;
    *step += 1;
    acc[8] = 0.0;
}

// PPO.spectral_normalize_actors (:398-404): sigma = ||W||_2 by power iteration on W^T W, warm-started from the previous optimiser
// step's right singular vector; W *= c / sigma when sigma > c.  One CTA per weight matrix, all matrices in one launch.
struct SpecSeg { float* w; int rows, cols; float* v; double* sigma; };
struct SpecParams { SpecSeg seg[kMaxHiddenL + 1]; int n_seg; float lipschitz; int max_iter; const int* stop; };
__global__ void __launch_bounds__(512) spectral_kernel(const SpecParams p) {
    pdl_launch_dependents();
    pdl_wait();
    if (p.stop && *p.stop) return;
    const SpecSeg& S = p.seg[blockIdx.x];
    __shared__ float v[256], u[256], part[2][256];
    __shared__ float red[16];
    __shared__ float s_norm;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows = S.rows, cols = S.cols;
    const float* __restrict__ W = S.w;
    for (int c = tid; c < cols; c += 512) v[c] = S.v[c];
    __syncthreads();
    // block-wide sum of squares of a shared vector of length n -> s_norm
    auto sumsq = [&](const float* x, int n) {
        float a = 0.f;
        for (int k = tid; k < n; k += 512) a += x[k] * x[k];
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) red[warp] = a;
        __syncthreads();
        if (tid == 0) { float t = 0.f; for (int w = 0; w < 16; ++w) t += red[w]; s_norm = t; }
        __syncthreads();
    };
    float sigma = 0.f, prev = -1.f;
    for (int it = 0; it < p.max_iter; ++it) {
        // u = W v: a warp per row (two rows in flight), lanes along the row: coalesced reads of W from L2
        for (int r0 = warp * 2; r0 < rows; r0 += 32) {
            float a0 = 0.f, a1 = 0.f;
            const bool two = r0 + 1 < rows;
            for (int c = lane; c < cols; c += 32) {
                const float vc = v[c];
                a0 = fmaf(__ldg(W + (long long)r0 * cols + c), vc, a0);
                if (two) a1 = fmaf(__ldg(W + (long long)(r0 + 1) * cols + c), vc, a1);
            }
            for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xffffffffu, a0, o); a1 += __shfl_xor_sync(0xffffffffu, a1, o); }
            if (lane == 0) { u[r0] = a0; if (two) u[r0 + 1] = a1; }
        }
        __syncthreads();
        sumsq(u, rows);                                              // sigma^2 estimate = ||W v||^2 for ||v|| = 1
        sigma = sqrtf(s_norm);
        // v = W^T u: thread <-> (column, half of the rows): lanes read consecutive columns of one row, 8 loads in flight
        {
            const int c = tid & 255, half = tid >> 8;
            if (c < cols) {
                const int rb = half * ((rows + 1) >> 1), re = half ? rows : ((rows + 1) >> 1);
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
                int r = rb;
                for (; r + 3 < re; r += 4) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[q] = fmaf(__ldg(W + (long long)(r + q) * cols + c), u[r + q], acc[q]);
                }
                for (; r < re; ++r) acc[0] = fmaf(__ldg(W + (long long)r * cols + c), u[r], acc[0]);
                part[half][c] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
            }
        }
        __syncthreads();
        for (int c = tid; c < cols; c += 512) v[c] = part[0][c] + part[1][c];
        __syncthreads();
        sumsq(v, cols);
        const float inv = s_norm > 0.f ? rsqrtf(s_norm) : 0.f;
        for (int c = tid; c < cols; c += 512) v[c] *= inv;
        __syncthreads();
        if (it >= 2 && fabsf(sigma - prev) <= 1e-6f * sigma) break;
        prev = sigma;
    }
    for (int c = tid; c < cols; c += 512) S.v[c] = v[c];
    if (tid == 0) *S.sigma = (double)sigma;
    if (p.lipschitz > 0.0f && sigma > p.lipschitz) {
        const float sc = p.lipschitz / sigma;                           // param.data *= lipschitz_const / spectral_norm
        for (long long e = tid; e < (long long)rows * cols; e += 512) S.w[e] *= sc;
    }
}

// fp32 master weights -> bf16 operand copies: W [out][ld_w] (K = in) and W^T [in][ld_t] (K = out), zero padded
struct PackSeg { const float* w; int out, in; bf16* wb; int ld_w; bf16* wt; int ld_t; int rows_t; };
struct PackParams { PackSeg seg[2 * (kMaxHiddenL + 1)]; int n_seg; const int* stop; };
__global__ void __launch_bounds__(256) pack_kernel(const PackParams p) {
    pdl_launch_dependents();
    pdl_wait();
    if (p.stop && *p.stop) return;
    const PackSeg& S = p.seg[blockIdx.y];
    const long long nw = (long long)S.out * S.ld_w, ntr = (long long)S.rows_t * S.ld_t;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nw + ntr; e += (long long)gridDim.x * blockDim.x) {
        if (e < nw) {
            const int o = (int)(e / S.ld_w), i = (int)(e % S.ld_w);
            S.wb[e] = __float2bfloat16_rn(i < S.in ? S.w[(long long)o * S.in + i] : 0.0f);
        } else if (S.wt) {
            const long long f = e - nw;
            const int i = (int)(f / S.ld_t), o = (int)(f % S.ld_t);
            S.wt[f] = __float2bfloat16_rn((i < S.in && o < S.out) ? S.w[(long long)o * S.in + i] : 0.0f);
        }
    }
}
// LSTM: Wcat [256][96] = [W_hh | W_ih | b_hi b_lo | 0] and W_hh^T [64][256]
__global__ void __launch_bounds__(256) pack_lstm_kernel(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, int in_dim, int ld_u,
                                                        bf16* wcat, bf16* whh_t, const int* stop) {
    pdl_launch_dependents();
    pdl_wait();
    if (stop && *stop) return;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < kG * ld_u + kH * kG; e += gridDim.x * blockDim.x) {
        if (e < kG * ld_u) {
            const int r = e / ld_u, c = e % ld_u;
            float v = 0.0f;
            if (c < kH) v = w_hh[r * kH + c];
            else if (c < kH + in_dim) v = w_ih[r * in_dim + (c - kH)];
            else if (c < kH + in_dim + 2) {
                const float b = b_ih[r] + b_hh[r];
                const float hi = __bfloat162float(__float2bfloat16_rn(b));
                v = (c == kH + in_dim) ? hi : b - hi;
            }
            wcat[e] = __float2bfloat16_rn(v);
        } else {
            const int f = e - kG * ld_u;
            const int j = f / kG, r = f % kG;                             // W_hh^T[j][r] = W_hh[r][j]
            whh_t[f] = __float2bfloat16_rn(w_hh[r * kH + j]);
        }
    }
}

}  // namespace ppo
}  // namespace taco

using namespace taco::ppo;

// ---------------------------------------------------------------------------------------------- the trainer object
struct MlpNet {
    int L = 0;                        // linear layers
    int s[kMaxHiddenL + 2] = {0};     // sizes [in, h1, ..., out]
    long long w_off[kMaxHiddenL + 1], b_off[kMaxHiddenL + 1];      // offsets into the flat parameter vector
    bf16* wb[kMaxHiddenL + 1]; int ld_w[kMaxHiddenL + 1];          // [out_pad16][pad8(in)]
    bf16* wt[kMaxHiddenL + 1]; int ld_t[kMaxHiddenL + 1];          // [in][pad8(out) (>= 16)]
    bf16* x_bm[kMaxHiddenL + 1]; int ld_x[kMaxHiddenL + 1];        // activations X_0 .. X_{L-1}: [B][pad8(s_l)]
    uint32_t* relu_mask[kMaxHiddenL + 1];                          // l >= 1: [ceil(s_l / 32)][B] words, bit j = pre-activation (sample, 32 c + j) > 0
    bf16* dz_bm[kMaxHiddenL + 2];                                  // dZ_1 .. dZ_L (index l): [B][ld]
    int ld_dz[kMaxHiddenL + 2];
    float* partial[kMaxHiddenL + 1]; int splits[kMaxHiddenL + 1]; int ld_p[kMaxHiddenL + 1];
    // K-major maps (forward / dX GEMMs) and MN-major maps of the same batch-major tensors (dW = dZ^T X reads both transposed)
    CUtensorMap m_x_bm[kMaxHiddenL + 1], m_x_mn[kMaxHiddenL + 1], m_w[kMaxHiddenL + 1], m_wt[kMaxHiddenL + 1];
    CUtensorMap m_dz_bm[kMaxHiddenL + 2], m_dz_mn[kMaxHiddenL + 2];
    CUtensorMap m_partial[kMaxHiddenL + 1]; bool partial_tma[kMaxHiddenL + 1];   // split-K partials as [splits * out][ld_p] fp32 (TMA-stored when out % 128 == 0)
};

struct TacoPPO {
    int device = 0, num_sms = 148;
    TacoPPOCfg cfg;
    int B = 0, A = 0, seq = 0, sd = 0, ld_u = 0;
    long long n_params = 0;
    long long off_log_std = 0, off_wih = 0, off_whh = 0, off_bih = 0, off_bhh = 0;
    float *prm = nullptr, *grad = nullptr, *adam_m = nullptr, *adam_v = nullptr;
    MlpNet actor, critic;
    // LSTM
    bf16 *wcat = nullptr, *whh_t = nullptr, *u_bm = nullptr, *gates = nullptr, *dg_bm = nullptr;     // dg_bm: [seq*B][256], all time steps
    float *c_state = nullptr, *dh = nullptr, *dc = nullptr, *lstm_partial = nullptr, *colsum_partial = nullptr;
    int lstm_splits = 1;
    CUtensorMap m_u_bm, m_u_mn, m_wcat, m_whh_t, m_dg_bm, m_dg_mn, m_lstm_partial, m_dh;
    // heads / side arrays
    float *mean = nullptr, *value = nullptr, *g_act = nullptr, *g_logp = nullptr, *g_adv = nullptr, *g_ret = nullptr;
    // bookkeeping
    double* acc = nullptr; float* log = nullptr; int *n_logged = nullptr, *stop = nullptr, *step = nullptr;
    int max_log = 4096;
    float* spec_v[kMaxHiddenL + 1] = {nullptr}; double* spec_sigma = nullptr;
    bool spec_cold = true;
    // the spectral projection + re-pack of the ACTOR runs on a side stream under the next minibatch's gather and critic forward
    cudaStream_t side = nullptr;
    cudaEvent_t ev_adam = nullptr, ev_actor = nullptr;
    bool actor_pending = false;
    std::vector<void*> allocs;
};

static int ppo_fail(int code, const std::string& m) { return taco::fail(code, m); }

template <typename T>
static cudaError_t dev_alloc(TacoPPO* t, T** p, size_t count) {
    cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
    if (e == cudaSuccess) { e = cudaMemset(*p, 0, count * sizeof(T)); t->allocs.push_back((void*)*p); }
    return e;
}

// launch with programmatic stream serialization: the kernel may be scheduled while its predecessor in the stream drains; every
// kernel launched this way starts with pdl_wait() (gemm_tc.cuh)
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    TACO_LAUNCHED();
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// c: tensor map of the output when the epilogue stores through TMA (p.tma_store is set from it)
static int launch_gemm(TacoPPO* t, const CUtensorMap& a, const CUtensorMap& b, GemmParams p, cudaStream_t s, const CUtensorMap* c = nullptr) {
    const int kb_total = (p.k + BK - 1) / BK;
    if (p.splits < 1) p.splits = 1;
    if (p.splits > kb_total) p.splits = kb_total;
    p.kb_per_split = (kb_total + p.splits - 1) / p.splits;
    p.splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;       // no empty split
    p.stop = t->stop;
    const int total = ((p.m + BM - 1) / BM) * p.splits;
    const int grid = total < t->num_sms ? total : t->num_sms;
    p.tma_store = c != nullptr;
    return launch_pdl(gemm_tc_kernel, dim3(grid), dim3(kGemmThreads), kGemmSmem, s, a, b, c ? *c : a, p) == cudaSuccess ? TACO_OK : ppo_fail(TACO_E_CUDA, "gemm_tc_kernel launch failed");
}

static int splits_for(int m, long long k, int num_sms) {
    const int m_tiles = (m + BM - 1) / BM;
    int s = num_sms / m_tiles;
    const long long kb = (k + BK - 1) / BK;
    if (s > kb) s = (int)kb;
    return s < 1 ? 1 : s;
}

static int setup_net(TacoPPO* t, MlpNet& n, long long& off, bool actor) {
    const int B = t->B;
    for (int l = 0; l < n.L; ++l) {
        const int in = n.s[l], out = n.s[l + 1];
        n.w_off[l] = off; off += (long long)in * out;
        n.b_off[l] = off; off += out;
        n.ld_w[l] = pad8(in);
        n.ld_t[l] = pad8(out) < 16 ? 16 : pad8(out);
        if (dev_alloc(t, &n.wb[l], (size_t)pad16(out) * n.ld_w[l]) != cudaSuccess) return TACO_E_NOMEM;
        if (dev_alloc(t, &n.wt[l], (size_t)pad8(in) * n.ld_t[l]) != cudaSuccess) return TACO_E_NOMEM;
        n.ld_x[l] = pad8(in);
        if (dev_alloc(t, &n.x_bm[l], (size_t)B * n.ld_x[l]) != cudaSuccess) return TACO_E_NOMEM;
        n.relu_mask[l] = nullptr;
        if (l > 0 && dev_alloc(t, &n.relu_mask[l], (size_t)((in + 31) / 32) * B) != cudaSuccess) return TACO_E_NOMEM;
        const int lz = l + 1;
        n.ld_dz[lz] = pad16(out);
        if (dev_alloc(t, &n.dz_bm[lz], (size_t)B * n.ld_dz[lz]) != cudaSuccess) return TACO_E_NOMEM;
        n.splits[l] = splits_for(out, B, t->num_sms);
        n.ld_p[l] = pad16(in);
        if (dev_alloc(t, &n.partial[l], (size_t)n.splits[l] * out * n.ld_p[l]) != cudaSuccess) return TACO_E_NOMEM;
        bool ok = make_map(&n.m_x_bm[l], n.x_bm[l], B, in, n.ld_x[l], BM);
        ok = ok && make_map(&n.m_x_mn[l], n.x_bm[l], B, n.ld_x[l], n.ld_x[l], BK);           // B operand of the dW GEMM (MN-major: 64 x 64 boxes)
        ok = ok && make_map(&n.m_w[l], n.wb[l], out, in, n.ld_w[l], pad16(out));             // B operand of the forward GEMM
        ok = ok && make_map(&n.m_wt[l], n.wt[l], in, pad8(out) < 16 ? 16 : out, n.ld_t[l], pad16(in));   // B operand of the dX GEMM
        ok = ok && make_map(&n.m_dz_bm[lz], n.dz_bm[lz], B, n.ld_dz[lz], n.ld_dz[lz], BM);   // A operand of the dX GEMM
        ok = ok && make_map(&n.m_dz_mn[lz], n.dz_bm[lz], B, n.ld_dz[lz], n.ld_dz[lz], BK);   // A operand of the dW GEMM (MN-major)
        n.partial_tma[l] = out % BM == 0;
        if (n.partial_tma[l]) ok = ok && make_map_f32(&n.m_partial[l], n.partial[l], (long long)n.splits[l] * out, n.ld_p[l], n.ld_p[l]);
        if (!ok) return ppo_fail(TACO_E_CUDA, "cuTensorMapEncodeTiled failed");
    }
    (void)actor;
    return TACO_OK;
}

extern "C" {

int taco_ppo_create(int device, const TacoPPOCfg* cfg, TacoPPO** out) {
    if (!cfg || !out) return ppo_fail(TACO_E_INVALID, "taco_ppo_create: null argument");
    *out = nullptr;
    if (cfg->batch < 128 || cfg->batch % 128) return ppo_fail(TACO_E_INVALID, "taco_ppo_create: the minibatch size must be a positive multiple of 128");
    if (cfg->lstm_hidden != kH) return ppo_fail(TACO_E_INVALID, "taco_ppo_create: lstm_hidden must be 64");
    if (cfg->act_dim < 1 || cfg->act_dim > 4) return ppo_fail(TACO_E_INVALID, "taco_ppo_create: act_dim must be in [1, 4]");
    if (cfg->seq_len < 1 || cfg->seq_len > kSeqMax || cfg->state_dim < 1 || kH + cfg->state_dim + 2 > 128 || cfg->obs_dim < 1 || cfg->obs_dim > 256)
        return ppo_fail(TACO_E_INVALID, "taco_ppo_create: seq_len <= 8, state_dim <= 62, obs_dim <= 256");
    if (cfg->n_actor_hidden < 1 || cfg->n_actor_hidden > kMaxHiddenL || cfg->n_critic_hidden < 1 || cfg->n_critic_hidden > kMaxHiddenL)
        return ppo_fail(TACO_E_INVALID, "taco_ppo_create: 1..4 hidden layers per MLP");
    for (int l = 0; l < cfg->n_actor_hidden; ++l)
        if (cfg->actor_hidden[l] < 16 || cfg->actor_hidden[l] > 256 || cfg->actor_hidden[l] % 16) return ppo_fail(TACO_E_INVALID, "taco_ppo_create: hidden widths must be multiples of 16 in [16, 256]");
    for (int l = 0; l < cfg->n_critic_hidden; ++l)
        if (cfg->critic_hidden[l] < 16 || cfg->critic_hidden[l] > 256 || cfg->critic_hidden[l] % 16) return ppo_fail(TACO_E_INVALID, "taco_ppo_create: hidden widths must be multiples of 16 in [16, 256]");
    int ndev = 0;
    PPO_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return ppo_fail(TACO_E_INVALID, "taco_ppo_create: no such CUDA device");
    DevGuard guard(device);
    cudaDeviceProp prop;
    PPO_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return ppo_fail(TACO_E_INVALID, "taco_ppo_create: the native update needs an sm_100 device (tcgen05); there is no fallback");
    if (!encode_fn()) return ppo_fail(TACO_E_CUDA, "taco_ppo_create: cuTensorMapEncodeTiled is not available from the driver");
    TacoPPO* t = new (std::nothrow) TacoPPO();
    if (!t) return ppo_fail(TACO_E_NOMEM, "host allocation failed");
    t->device = device; t->cfg = *cfg; t->num_sms = prop.multiProcessorCount;
    t->B = cfg->batch; t->A = cfg->act_dim; t->seq = cfg->seq_len; t->sd = cfg->state_dim;
    t->ld_u = pad8(kH + cfg->state_dim + 2);
    const int B = t->B;
    int rc = TACO_OK;
    auto bail = [&](int code, const char* msg) { taco_ppo_destroy(t); return ppo_fail(code, msg); };
    // ---- flat parameter layout: log_std | actor MLP (W, b per layer) | LSTM W_ih, W_hh, b_ih, b_hh | critic MLP
    long long off = 0;
    t->off_log_std = off; off += t->A;
    t->actor.L = cfg->n_actor_hidden + 1;
    t->actor.s[0] = cfg->obs_dim;
    for (int l = 0; l < cfg->n_actor_hidden; ++l) t->actor.s[l + 1] = cfg->actor_hidden[l];
    t->actor.s[t->actor.L] = t->A;
    if ((rc = setup_net(t, t->actor, off, true)) != TACO_OK) return bail(rc, "taco_ppo_create: actor buffers");
    t->off_wih = off; off += (long long)kG * t->sd;
    t->off_whh = off; off += (long long)kG * kH;
    t->off_bih = off; off += kG;
    t->off_bhh = off; off += kG;
    t->critic.L = cfg->n_critic_hidden + 1;
    t->critic.s[0] = kH;
    for (int l = 0; l < cfg->n_critic_hidden; ++l) t->critic.s[l + 1] = cfg->critic_hidden[l];
    t->critic.s[t->critic.L] = 1;
    if ((rc = setup_net(t, t->critic, off, false)) != TACO_OK) return bail(rc, "taco_ppo_create: critic buffers");
    t->n_params = off;
    cudaError_t ce = dev_alloc(t, &t->prm, (size_t)off);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->grad, (size_t)off);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->adam_m, (size_t)off);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->adam_v, (size_t)off);
    // ---- LSTM buffers
    const long long SB = (long long)t->seq * B;
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->wcat, (size_t)kG * t->ld_u);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->whh_t, (size_t)kH * kG);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->u_bm, (size_t)SB * t->ld_u);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->gates, (size_t)SB * kG);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->dg_bm, (size_t)SB * kG);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->colsum_partial, (size_t)2 * (kMaxHiddenL + 1) * ((B + kColRows - 1) / kColRows) * kColLd);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->c_state, (size_t)SB * kH);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->dh, (size_t)B * kH);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->dc, (size_t)B * kH);
    t->lstm_splits = splits_for(kG, SB, t->num_sms);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->lstm_partial, (size_t)t->lstm_splits * kG * t->ld_u);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->mean, (size_t)B * t->A);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->value, (size_t)B);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->g_act, (size_t)B * t->A);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->g_logp, (size_t)B);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->g_adv, (size_t)B);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->g_ret, (size_t)B);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->acc, (size_t)kNumStat);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->log, (size_t)t->max_log * kLogCols);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->n_logged, 1);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->stop, 1);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->step, 1);
    if (ce == cudaSuccess) ce = dev_alloc(t, &t->spec_sigma, (size_t)(kMaxHiddenL + 1));
    for (int l = 0; l < t->actor.L && ce == cudaSuccess; ++l) {
        ce = dev_alloc(t, &t->spec_v[l], (size_t)t->actor.s[l]);
        if (ce == cudaSuccess) {
            std::vector<float> ones((size_t)t->actor.s[l], 1.0f / std::sqrt((float)t->actor.s[l]));
            ce = cudaMemcpy(t->spec_v[l], ones.data(), ones.size() * sizeof(float), cudaMemcpyHostToDevice);
        }
    }
    if (ce != cudaSuccess) return bail(ce == cudaErrorMemoryAllocation ? TACO_E_NOMEM : TACO_E_CUDA, "taco_ppo_create: device allocation failed");
    bool ok = make_map(&t->m_u_bm, t->u_bm, SB, t->ld_u, t->ld_u, BM);
    ok = ok && make_map(&t->m_u_mn, t->u_bm, SB, t->ld_u, t->ld_u, BK);
    ok = ok && make_map(&t->m_wcat, t->wcat, kG, t->ld_u, t->ld_u, kG);
    ok = ok && make_map(&t->m_whh_t, t->whh_t, kH, kG, kG, kH);
    ok = ok && make_map(&t->m_dg_bm, t->dg_bm, SB, kG, kG, BM);
    ok = ok && make_map(&t->m_dg_mn, t->dg_bm, SB, kG, kG, BK);
    ok = ok && make_map_f32(&t->m_lstm_partial, t->lstm_partial, (long long)t->lstm_splits * kG, t->ld_u, t->ld_u);
    ok = ok && make_map_f32(&t->m_dh, t->dh, B, kH, kH);
    if (!ok) return bail(TACO_E_CUDA, "taco_ppo_create: cuTensorMapEncodeTiled failed");
    if (cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem) != cudaSuccess)
        return bail(TACO_E_CUDA, "taco_ppo_create: cudaFuncSetAttribute failed");
    if (cudaStreamCreateWithFlags(&t->side, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&t->ev_adam, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&t->ev_actor, cudaEventDisableTiming) != cudaSuccess)
        return bail(TACO_E_CUDA, "taco_ppo_create: stream / event creation failed");
    *out = t;
    return TACO_OK;
}

int taco_ppo_destroy(TacoPPO* t) {
    if (!t) return TACO_OK;
    DevGuard guard(t->device);
    for (void* p : t->allocs) cudaFree(p);
    if (t->side) cudaStreamDestroy(t->side);
    if (t->ev_adam) cudaEventDestroy(t->ev_adam);
    if (t->ev_actor) cudaEventDestroy(t->ev_actor);
    delete t;
    return TACO_OK;
}

int taco_ppo_num_params(TacoPPO* t, int64_t* n) {
    if (!t || !n) return ppo_fail(TACO_E_INVALID, "taco_ppo_num_params: null argument");
    *n = t->n_params;
    return TACO_OK;
}

// offsets of the parameter tensors inside the flat vector, in this order: log_std; actor (W, b) x layers; LSTM weight_ih_l0,
// weight_hh_l0, bias_ih_l0, bias_hh_l0; critic (W, b) x layers.  `count` = capacity of `offsets`; returns the number written.
int taco_ppo_param_offsets(TacoPPO* t, int64_t* offsets, int32_t count, int32_t* n_out) {
    if (!t || !offsets || !n_out) return ppo_fail(TACO_E_INVALID, "taco_ppo_param_offsets: null argument");
    std::vector<int64_t> v;
    v.push_back(t->off_log_std);
    for (int l = 0; l < t->actor.L; ++l) { v.push_back(t->actor.w_off[l]); v.push_back(t->actor.b_off[l]); }
    v.push_back(t->off_wih); v.push_back(t->off_whh); v.push_back(t->off_bih); v.push_back(t->off_bhh);
    for (int l = 0; l < t->critic.L; ++l) { v.push_back(t->critic.w_off[l]); v.push_back(t->critic.b_off[l]); }
    if ((int)v.size() > count) return ppo_fail(TACO_E_INVALID, "taco_ppo_param_offsets: buffer too small");
    for (size_t i = 0; i < v.size(); ++i) offsets[i] = v[i];
    *n_out = (int32_t)v.size();
    return TACO_OK;
}

// device addresses of the flat fp32 vectors (each n_params floats): parameters, gradient of the last backward, Adam m, Adam v;
// and of the int32 optimiser step counter.  Any pointer may be NULL.
int taco_ppo_buffers(TacoPPO* t, float** params, float** grad, float** adam_m, float** adam_v, int32_t** step) {
    if (!t) return ppo_fail(TACO_E_INVALID, "taco_ppo_buffers: null argument");
    if (params) *params = t->prm;
    if (grad) *grad = t->grad;
    if (adam_m) *adam_m = t->adam_m;
    if (adam_v) *adam_v = t->adam_v;
    if (step) *step = t->step;
    return TACO_OK;
}

// the caller's stream waits for the actor's projection / re-pack still running on the side stream
static int join_actor(TacoPPO* t, cudaStream_t s) {
    if (t->actor_pending) {
        PPO_CUDA(cudaStreamWaitEvent(s, t->ev_actor, 0));
        t->actor_pending = false;
    }
    return TACO_OK;
}

// which: 0 = everything, 1 = the actor's matrices only, 2 = critic MLP + LSTM only
static int repack(TacoPPO* t, cudaStream_t s, bool honour_stop, int which = 0) {
    PackParams pp;
    memset(&pp, 0, sizeof(pp));
    for (int k = 0; k < 2; ++k) {
        if ((which == 1 && k == 1) || (which == 2 && k == 0)) continue;
        MlpNet& n = k == 0 ? t->actor : t->critic;
        for (int l = 0; l < n.L; ++l) {
            PackSeg& S = pp.seg[pp.n_seg++];
            S.w = t->prm + n.w_off[l]; S.out = n.s[l + 1]; S.in = n.s[l];
            S.wb = n.wb[l]; S.ld_w = n.ld_w[l]; S.wt = n.wt[l]; S.ld_t = n.ld_t[l]; S.rows_t = pad8(n.s[l]);
        }
    }
    pp.stop = honour_stop ? t->stop : nullptr;
    launch_pdl(pack_kernel, dim3(32, pp.n_seg), dim3(256), 0, s, pp);
    if (which != 1) {
        launch_pdl(pack_lstm_kernel, dim3(64), dim3(256), 0, s, (const float*)(t->prm + t->off_wih), (const float*)(t->prm + t->off_whh), (const float*)(t->prm + t->off_bih),
                   (const float*)(t->prm + t->off_bhh), t->sd, t->ld_u, t->wcat, t->whh_t, (const int*)(honour_stop ? t->stop : nullptr));
    }
    return cudaGetLastError() == cudaSuccess ? TACO_OK : ppo_fail(TACO_E_CUDA, "pack kernels failed");
}

// call after writing the parameter vector (taco_ppo_buffers) from outside: refreshes the bf16 operand copies
int taco_ppo_params_changed(TacoPPO* t, void* stream) {
    if (!t) return ppo_fail(TACO_E_INVALID, "taco_ppo_params_changed: null argument");
    DevGuard guard(t->device);
    const int rcj = join_actor(t, (cudaStream_t)stream);
    if (rcj != TACO_OK) return rcj;
    return repack(t, (cudaStream_t)stream, false);
}

// start of one PPO.update: clears the early-stop flag and the log
int taco_ppo_begin_update(TacoPPO* t, void* stream) {
    if (!t) return ppo_fail(TACO_E_INVALID, "taco_ppo_begin_update: null argument");
    DevGuard guard(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    PPO_CUDA(cudaMemsetAsync(t->stop, 0, sizeof(int), s));
    PPO_CUDA(cudaMemsetAsync(t->n_logged, 0, sizeof(int), s));
    PPO_CUDA(cudaMemsetAsync(t->acc, 0, kNumStat * sizeof(double), s));
    return TACO_OK;
}

static Hyper to_hyper(const TacoPPOHyper* h) {
    Hyper r;
    r.lr = h->lr; r.clip = h->clip; r.target_kl = h->target_kl; r.max_grad = h->max_grad; r.pi_coef = h->pi_coef; r.vf_coef = h->vf_coef;
    r.ent_coef = h->ent_coef; r.lipschitz = h->lipschitz; r.use_lipschitz = h->use_lipschitz; r.world = h->world < 1 ? 1 : h->world;
    return r;
}

static int mlp_forward(TacoPPO* t, MlpNet& n, bool actor, cudaStream_t s) {
    for (int l = 0; l < n.L; ++l) {
        GemmParams p;
        memset(&p, 0, sizeof(p));
        p.m = t->B; p.n = n.s[l + 1]; p.k = n.s[l]; p.n_tile = pad16(n.s[l + 1]); p.splits = 1;
        p.bias = t->prm + n.b_off[l];
        if (l + 1 < n.L) {
            p.epi = EPI_BIAS_RELU; p.n_valid = pad8(n.s[l + 1]);
            p.mask_out = n.relu_mask[l + 1]; p.ld_mask = t->B;
        } else if (actor) {
            p.epi = EPI_TANH_F32; p.n_valid = t->A; p.out_f32 = t->mean; p.ldc = t->A;
        } else {
            p.epi = EPI_F32; p.n_valid = 1; p.out_f32 = t->value; p.ldc = 1;
        }
        const int rc = launch_gemm(t, n.m_x_bm[l], n.m_w[l], p, s, l + 1 < n.L ? &n.m_x_bm[l + 1] : nullptr);   // X_{l+1} leaves through its own (operand) map
        if (rc != TACO_OK) return rc;
    }
    return TACO_OK;
}

// backward of an MLP whose dZ_L (both layouts) is already written.  dx_out != null: also d loss / d input as fp32 [B][s0] (critic -> LSTM)
static int mlp_backward(TacoPPO* t, MlpNet& n, float* dx_out, cudaStream_t s) {
    for (int l = n.L - 1; l >= 0; --l) {
        const int lz = l + 1, in = n.s[l], out = n.s[l + 1];
        {   // dW_l = dZ^T X_{l-1}: M = out, N = in, K = batch
            GemmParams p;
            memset(&p, 0, sizeof(p));
            p.m = out; p.n = in; p.k = t->B; p.n_tile = pad16(in); p.splits = n.splits[l];
            p.epi = EPI_F32; p.n_valid = pad16(in);
            p.out_f32 = n.partial[l]; p.ldc = n.ld_p[l]; p.split_stride = (long long)out * n.ld_p[l];
            p.mn_major = 1;
            const int rc = launch_gemm(t, n.m_dz_mn[lz], n.m_x_mn[l], p, s, n.partial_tma[l] ? &n.m_partial[l] : nullptr);
            if (rc != TACO_OK) return rc;
        }
        if (l > 0 || dx_out) {   // dX_{l-1} = dZ W_l: M = batch, N = in, K = out
            GemmParams p;
            memset(&p, 0, sizeof(p));
            p.m = t->B; p.n = in; p.k = n.ld_dz[lz] < n.ld_t[l] ? n.ld_dz[lz] : n.ld_t[l]; p.n_tile = pad16(in); p.splits = 1;
            if (l > 0) {
                p.epi = EPI_RELUBWD; p.n_valid = pad8(in);
                p.mask_in = n.relu_mask[l]; p.ld_mask = t->B;
            } else {
                p.epi = EPI_F32; p.n_valid = in; p.out_f32 = dx_out; p.ldc = in;
            }
            const int rc = launch_gemm(t, n.m_dz_bm[lz], n.m_wt[l], p, s, l > 0 ? &n.m_dz_bm[l] : (dx_out == t->dh ? &t->m_dh : nullptr));
            if (rc != TACO_OK) return rc;
        }
    }
    return TACO_OK;
}

// phase 1: gather + forward + loss.  The rollout tensors are flat views [N_total][...] on the device; idx_dev = B int64 row indices.
int taco_ppo_forward_loss(TacoPPO* t, const TacoPPOHyper* hyper, const float* obs, const float* states, const float* act, const float* old_logp,
                          const float* adv, const float* ret, const int64_t* idx_dev, void* stream) {
    if (!t || !hyper || !obs || !states || !act || !old_logp || !adv || !ret || !idx_dev) return ppo_fail(TACO_E_INVALID, "taco_ppo_forward_loss: null argument");
    DevGuard guard(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int B = t->B;
    GatherParams g;
    memset(&g, 0, sizeof(g));
    g.obs = obs; g.obs_dim = t->cfg.obs_dim; g.states = states; g.seq = t->seq; g.state_dim = t->sd;
    g.act = act; g.logp = old_logp; g.adv = adv; g.ret = ret; g.idx = (const long long*)idx_dev; g.B = B; g.act_dim = t->A;
    g.x0_bm = t->actor.x_bm[0]; g.ld_x0 = t->actor.ld_x[0];
    g.u_bm = t->u_bm; g.ld_u = t->ld_u;
    g.g_act = t->g_act; g.g_logp = t->g_logp; g.g_adv = t->g_adv; g.g_ret = t->g_ret; g.stop = t->stop;
    launch_pdl(gather_kernel, dim3((B + 31) / 32), dim3(256), 0, s, g);
    int rc = TACO_OK;
    // critic first: seq LSTM steps, each one GEMM [h_{t-1} | x_t | 1 1] Wcat^T with the gate math in the epilogue, then its MLP.  The
    // previous optimiser step's projection + re-pack of the ACTOR may still be running on the side stream under these launches.
    for (int k = 0; k < t->seq; ++k) {
        GemmParams p;
        memset(&p, 0, sizeof(p));
        p.m = B; p.n = kG; p.k = t->ld_u; p.n_tile = kG; p.splits = 1; p.a_row0 = k * B; p.epi = EPI_LSTM; p.n_valid = kG;
        p.lstm.c_prev = k > 0 ? t->c_state + (size_t)(k - 1) * B * kH : nullptr;
        p.lstm.c_out = t->c_state + (size_t)k * B * kH;
        p.lstm.gates_out = t->gates + (size_t)k * B * kG;
        if (k + 1 < t->seq) {
            p.lstm.h_bm = t->u_bm + (size_t)(k + 1) * B * t->ld_u; p.lstm.ld_h_bm = t->ld_u;
        } else {
            p.lstm.h_bm = t->critic.x_bm[0]; p.lstm.ld_h_bm = t->critic.ld_x[0];
        }
        rc = launch_gemm(t, t->m_u_bm, t->m_wcat, p, s);
        if (rc != TACO_OK) return rc;
    }
    rc = mlp_forward(t, t->critic, false, s);
    if (rc != TACO_OK) return rc;
    rc = join_actor(t, s);
    if (rc != TACO_OK) return rc;
    rc = mlp_forward(t, t->actor, true, s);
    if (rc != TACO_OK) return rc;
    LossParams L;
    memset(&L, 0, sizeof(L));
    L.mean = t->mean; L.value = t->value; L.act = t->g_act; L.old_logp = t->g_logp; L.adv = t->g_adv; L.ret = t->g_ret;
    L.log_std = t->prm + t->off_log_std; L.B = B; L.act_dim = t->A; L.h = to_hyper(hyper);
    L.dz_bm = t->actor.dz_bm[t->actor.L];
    L.dv_bm = t->critic.dz_bm[t->critic.L];
    L.acc = t->acc; L.stop = t->stop;
    launch_pdl(loss_kernel, dim3((B + 255) / 256), dim3(256), 0, s, L);
    PPO_CUDA(cudaGetLastError());
    return TACO_OK;
}

// device address of the 8 float64 loss accumulators [sum surrogate, sum value, sum kl, -, d log_std x 4]: a data-parallel job
// all-reduces (SUM) them between taco_ppo_forward_loss and taco_ppo_decide and passes world > 1 in the hyper-parameters
int taco_ppo_loss_sums(TacoPPO* t, double** acc_dev) {
    if (!t || !acc_dev) return ppo_fail(TACO_E_INVALID, "taco_ppo_loss_sums: null argument");
    *acc_dev = t->acc;
    return TACO_OK;
}

// phase 2: log row + KL early-stop decision (device side)
int taco_ppo_decide(TacoPPO* t, const TacoPPOHyper* hyper, void* stream) {
    if (!t || !hyper) return ppo_fail(TACO_E_INVALID, "taco_ppo_decide: null argument");
    DevGuard guard(t->device);
    DecideParams d;
    memset(&d, 0, sizeof(d));
    d.acc = t->acc; d.log = t->log; d.n_logged = t->n_logged; d.stop = t->stop; d.max_log = t->max_log;
    d.log_std = t->prm + t->off_log_std; d.act_dim = t->A; d.B = t->B; d.h = to_hyper(hyper);
    d.grad_log_std = t->grad + t->off_log_std;
    launch_pdl(decide_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, d);
    PPO_CUDA(cudaGetLastError());
    return TACO_OK;
}

// phase 3: backward -> flat gradient (local; a data-parallel job all-reduces (SUM) it before taco_ppo_apply)
int taco_ppo_backward(TacoPPO* t, void* stream) {
    if (!t) return ppo_fail(TACO_E_INVALID, "taco_ppo_backward: null argument");
    DevGuard guard(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int B = t->B;
    int rc = mlp_backward(t, t->actor, nullptr, s);
    if (rc != TACO_OK) return rc;
    rc = mlp_backward(t, t->critic, t->dh, s);
    if (rc != TACO_OK) return rc;
    const long long SB = (long long)t->seq * B;
    for (int k = t->seq - 1; k >= 0; --k) {
        LstmBwdParams b;
        memset(&b, 0, sizeof(b));
        b.dh = t->dh; b.dc = t->dc; b.gates = t->gates + (size_t)k * B * kG;
        b.c_prev = k > 0 ? t->c_state + (size_t)(k - 1) * B * kH : nullptr; b.c_cur = t->c_state + (size_t)k * B * kH;
        b.dg_bm = t->dg_bm + (size_t)k * B * kG; b.B = B; b.first = (k == t->seq - 1); b.stop = t->stop;
        launch_pdl(lstm_bwd_kernel, dim3((B + 3) / 4), dim3(256), 0, s, b);
        if (k > 0) {   // dh_{t-1} = dgates_t W_hh: M = batch, N = 64, K = 256
            GemmParams p;
            memset(&p, 0, sizeof(p));
            p.m = B; p.n = kH; p.k = kG; p.n_tile = kH; p.splits = 1; p.a_row0 = k * B; p.epi = EPI_F32; p.n_valid = kH; p.out_f32 = t->dh; p.ldc = kH;
            rc = launch_gemm(t, t->m_dg_bm, t->m_whh_t, p, s, &t->m_dh);
            if (rc != TACO_OK) return rc;
        }
    }
    {   // dWcat = dgates^T [h | x | 1 1] over all time steps: M = 256, N = ld_u, K = seq * batch (both operands read transposed)
        GemmParams p;
        memset(&p, 0, sizeof(p));
        p.m = kG; p.n = t->ld_u; p.k = (int)SB; p.n_tile = pad16(t->ld_u); p.splits = t->lstm_splits; p.epi = EPI_F32; p.n_valid = pad16(t->ld_u);
        p.out_f32 = t->lstm_partial; p.ldc = t->ld_u; p.split_stride = (long long)kG * t->ld_u;
        if (p.n_valid > t->ld_u) p.n_valid = t->ld_u;
        p.mn_major = 1;
        rc = launch_gemm(t, t->m_dg_mn, t->m_u_mn, p, s, &t->m_lstm_partial);
        if (rc != TACO_OK) return rc;
    }
    // bias gradients = column sums of the batch-major dZ: partial rows here, summed by the assembly kernel below
    ColSumParams cs;
    memset(&cs, 0, sizeof(cs));
    const int col_chunks = (B + kColRows - 1) / kColRows;
    for (int k = 0; k < 2; ++k) {
        MlpNet& n = k == 0 ? t->actor : t->critic;
        for (int l = 0; l < n.L; ++l) {
            ColSumSeg& S = cs.seg[cs.n_seg++];
            S.src = n.dz_bm[l + 1]; S.ld = n.ld_dz[l + 1];
        }
    }
    cs.B = B; cs.partial = t->colsum_partial; cs.stop = t->stop;
    launch_pdl(colsum_kernel, dim3(col_chunks, cs.n_seg), dim3(256), 0, s, cs);
    // weight gradients from the split-K partials
    GradParams gp;
    memset(&gp, 0, sizeof(gp));
    long long elems = 0;
    auto add = [&](const float* partial, int splits, long long stride, int ld, int col0, int rows, int cols, long long dst) {
        GradSeg& S = gp.seg[gp.n_seg++];
        S.partial = partial; S.splits = splits; S.split_stride = stride; S.ld = ld; S.col0 = col0; S.rows = rows; S.cols = cols; S.dst = dst;
        elems += (long long)rows * cols;
    };
    for (int k = 0; k < 2; ++k) {
        MlpNet& n = k == 0 ? t->actor : t->critic;
        for (int l = 0; l < n.L; ++l) {
            // the launch may have merged splits: recompute the number actually written
            const int kb_total = (B + BK - 1) / BK;
            int sp = n.splits[l] > kb_total ? kb_total : n.splits[l];
            const int per = (kb_total + sp - 1) / sp;
            sp = (kb_total + per - 1) / per;
            add(n.partial[l], sp, (long long)n.s[l + 1] * n.ld_p[l], n.ld_p[l], 0, n.s[l + 1], n.s[l], n.w_off[l]);
            add(t->colsum_partial + (long long)(k * t->actor.L + l) * col_chunks * kColLd, col_chunks, kColLd, kColLd, 0, 1, n.s[l + 1], n.b_off[l]);
        }
    }
    {
        const int kb_total = (int)((SB + BK - 1) / BK);
        int sp = t->lstm_splits > kb_total ? kb_total : t->lstm_splits;
        const int per = (kb_total + sp - 1) / sp;
        sp = (kb_total + per - 1) / per;
        const long long stride = (long long)kG * t->ld_u;
        add(t->lstm_partial, sp, stride, t->ld_u, 0, kG, kH, t->off_whh);
        add(t->lstm_partial, sp, stride, t->ld_u, kH, kG, t->sd, t->off_wih);
        add(t->lstm_partial, sp, stride, t->ld_u, kH + t->sd, kG, 1, t->off_bih);
        add(t->lstm_partial, sp, stride, t->ld_u, kH + t->sd, kG, 1, t->off_bhh);
    }
    gp.grad = t->grad; gp.acc = t->acc; gp.stop = t->stop;
    launch_pdl(grad_assemble_kernel, dim3((unsigned)((elems + 255) / 256)), dim3(256), 0, s, gp);
    PPO_CUDA(cudaGetLastError());
    return TACO_OK;
}

// phase 4: clip_grad_norm_ + Adam + spectral projection + re-pack of the bf16 operand copies
int taco_ppo_apply(TacoPPO* t, const TacoPPOHyper* hyper, void* stream) {
    if (!t || !hyper) return ppo_fail(TACO_E_INVALID, "taco_ppo_apply: null argument");
    DevGuard guard(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    const Hyper h = to_hyper(hyper);
    launch_pdl(grad_norm_kernel, dim3(64), dim3(256), 0, s, (const float*)t->grad, t->n_params, 1.0f / (float)h.world, t->acc, (const int*)t->stop);
    launch_pdl(adam_kernel, dim3(t->num_sms), dim3(256), 0, s, t->prm, t->adam_m, t->adam_v, (const float*)t->grad, t->n_params, h, t->step, t->acc, t->log, (const int*)t->n_logged,
               t->max_log, (const int*)t->stop);
    launch_pdl(finish_step_kernel, dim3(1), dim3(1), 0, s, t->step, t->acc, (const int*)t->stop);
    if (h.use_lipschitz) {
        SpecParams sp;
        memset(&sp, 0, sizeof(sp));
        for (int l = 0; l < t->actor.L; ++l) {
            SpecSeg& S = sp.seg[sp.n_seg++];
            S.w = t->prm + t->actor.w_off[l]; S.rows = t->actor.s[l + 1]; S.cols = t->actor.s[l]; S.v = t->spec_v[l]; S.sigma = t->spec_sigma + l;
        }
        sp.lipschitz = h.lipschitz; sp.max_iter = t->spec_cold ? 4000 : 200; sp.stop = t->stop;
        t->spec_cold = false;
        // actor: projection, then the bf16 copies of the projected matrices, on the side stream (4 CTAs of latency-bound power
        // iteration); the caller's stream goes on with the critic's copies and the next minibatch until it needs the actor
        PPO_CUDA(cudaEventRecord(t->ev_adam, s));
        PPO_CUDA(cudaStreamWaitEvent(t->side, t->ev_adam, 0));
        spectral_kernel<<<sp.n_seg, 512, 0, t->side>>>(sp); TACO_LAUNCHED();
        PPO_CUDA(cudaGetLastError());
        int rc = repack(t, t->side, true, 1);
        if (rc != TACO_OK) return rc;
        PPO_CUDA(cudaEventRecord(t->ev_actor, t->side));
        t->actor_pending = true;
        return repack(t, s, true, 2);
    }
    PPO_CUDA(cudaGetLastError());
    return repack(t, s, true);
}

// end of an update: synchronises and returns the log rows (n_rows x 8 floats: policy-gradient loss, value loss, entropy loss,
// total loss, approx KL, gradient norm, stopped-here flag, clip coefficient), the number of optimiser steps taken overall and
// whether the KL early stop fired.  log_host may be NULL.
int taco_ppo_end_update(TacoPPO* t, float* log_host, int32_t max_rows, int32_t* n_rows, int32_t* optim_steps, int32_t* early_stop, void* stream) {
    if (!t) return ppo_fail(TACO_E_INVALID, "taco_ppo_end_update: null argument");
    DevGuard guard(t->device);
    cudaStream_t s = (cudaStream_t)stream;
    {
        const int rcj = join_actor(t, s);
        if (rcj != TACO_OK) return rcj;
    }
    int h[3] = {0, 0, 0};
    PPO_CUDA(cudaMemcpyAsync(&h[0], t->n_logged, sizeof(int), cudaMemcpyDeviceToHost, s));
    PPO_CUDA(cudaMemcpyAsync(&h[1], t->step, sizeof(int), cudaMemcpyDeviceToHost, s));
    PPO_CUDA(cudaMemcpyAsync(&h[2], t->stop, sizeof(int), cudaMemcpyDeviceToHost, s));
    PPO_CUDA(cudaStreamSynchronize(s));
    int rows = h[0] < t->max_log ? h[0] : t->max_log;
    if (rows > max_rows) rows = max_rows;
    if (log_host && rows > 0) PPO_CUDA(cudaMemcpy(log_host, t->log, (size_t)rows * kLogCols * sizeof(float), cudaMemcpyDeviceToHost));
    if (n_rows) *n_rows = rows;
    if (optim_steps) *optim_steps = h[1];
    if (early_stop) *early_stop = h[2];
    return TACO_OK;
}

// sigma of every actor weight matrix as measured by the last projection (n_actor_layers doubles, host)
int taco_ppo_sigmas(TacoPPO* t, double* out_host) {
    if (!t || !out_host) return ppo_fail(TACO_E_INVALID, "taco_ppo_sigmas: null argument");
    DevGuard guard(t->device);
    PPO_CUDA(cudaDeviceSynchronize());
    PPO_CUDA(cudaMemcpy(out_host, t->spec_sigma, (size_t)t->actor.L * sizeof(double), cudaMemcpyDeviceToHost));
    return TACO_OK;
}

// test hook: device addresses of the forward results of the last taco_ppo_forward_loss: action mean (batch, act_dim) and value (batch)
int taco_ppo_debug_outputs(TacoPPO* t, float** mean_dev, float** value_dev) {
    if (!t) return ppo_fail(TACO_E_INVALID, "taco_ppo_debug_outputs: null argument");
    if (mean_dev) *mean_dev = t->mean;
    if (value_dev) *value_dev = t->value;
    return TACO_OK;
}

// test hook: D[M][N] (fp32, ld = N) = A[M][K] B[N][K]^T for bf16 row-major device matrices (ld = K, K multiple of 8) through
// gemm_tc_kernel with `splits` split-K partials summed on the host side of the call
static int gemm_selftest_impl(int device, const void* a_bf16, const void* b_bf16, float* d_f32, int32_t m, int32_t n, int32_t k, int32_t splits, bool mn, void* stream) {
    if (!a_bf16 || !b_bf16 || !d_f32 || m < 1 || n < 1 || n > 256 || k < 8 || (k & 7)) return ppo_fail(TACO_E_INVALID, "taco_gemm_selftest: bad argument");
    DevGuard guard(device);
    cudaStream_t s = (cudaStream_t)stream;
    if (!encode_fn()) return ppo_fail(TACO_E_CUDA, "cuTensorMapEncodeTiled unavailable");
    CUtensorMap ma, mb;
    const bool ok = mn ? make_map(&ma, (const bf16*)a_bf16, k, m, m, BK) && make_map(&mb, (const bf16*)b_bf16, k, n, n, BK)
                       : make_map(&ma, (const bf16*)a_bf16, m, k, k, BM) && make_map(&mb, (const bf16*)b_bf16, n, k, k, pad16(n));
    if (!ok) return ppo_fail(TACO_E_CUDA, "tensor map");
    PPO_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
    const int kb_total = (k + BK - 1) / BK;
    int sp = splits < 1 ? 1 : (splits > kb_total ? kb_total : splits);
    const int per = (kb_total + sp - 1) / sp;
    sp = (kb_total + per - 1) / per;
    float* part = nullptr;
    PPO_CUDA(cudaMalloc(&part, (size_t)sp * m * n * sizeof(float)));
    // (compute-sanitizer's initcheck does not observe writes made by cp.async.bulk.tensor stores: without this the partials the
    // TMA-store epilogue writes read as uninitialised in the assembly kernel below)
    PPO_CUDA(cudaMemsetAsync(part, 0, (size_t)sp * m * n * sizeof(float), s));
    GemmParams p;
    memset(&p, 0, sizeof(p));
    p.m = m; p.n = n; p.k = k; p.n_tile = pad16(n); p.splits = sp; p.kb_per_split = per; p.epi = EPI_F32; p.n_valid = n; p.mn_major = mn ? 1 : 0;
    p.out_f32 = part; p.ldc = n; p.split_stride = (long long)m * n;
    cudaDeviceProp prop;
    PPO_CUDA(cudaGetDeviceProperties(&prop, device));
    const int total = ((m + BM - 1) / BM) * sp;
    CUtensorMap mc = ma;
    p.tma_store = (m % BM == 0 && n % 4 == 0) ? 1 : 0;      // exercise the TMA-store epilogue where its shape rule holds
    if (p.tma_store && !make_map_f32(&mc, part, (long long)sp * m, n, n)) return ppo_fail(TACO_E_CUDA, "tensor map");
    gemm_tc_kernel<<<total < prop.multiProcessorCount ? total : prop.multiProcessorCount, kGemmThreads, kGemmSmem, s>>>(ma, mb, mc, p); TACO_LAUNCHED();
    GradParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.n_seg = 1;
    gp.seg[0].partial = part; gp.seg[0].splits = sp; gp.seg[0].split_stride = (long long)m * n; gp.seg[0].ld = n; gp.seg[0].rows = m; gp.seg[0].cols = n;
    gp.grad = d_f32;
    grad_assemble_kernel<<<(unsigned)(((long long)m * n + 255) / 256), 256, 0, s>>>(gp); TACO_LAUNCHED();
    const cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(part);
    if (e != cudaSuccess) return ppo_fail(TACO_E_CUDA, std::string("taco_gemm_selftest: ") + cudaGetErrorString(e));
    return TACO_OK;
}

// D[m][n] = A[m][k] B[n][k]^T with both operands K-major (k contiguous)
int taco_gemm_selftest(int device, const void* a_bf16, const void* b_bf16, float* d_f32, int32_t m, int32_t n, int32_t k, int32_t splits, void* stream) {
    return gemm_selftest_impl(device, a_bf16, b_bf16, d_f32, m, n, k, splits, false, stream);
}
// the same product from TRANSPOSED operands: at[k][m], bt[k][n] (m, n multiples of 8)
int taco_gemm_selftest_mn(int device, const void* at_bf16, const void* bt_bf16, float* d_f32, int32_t m, int32_t n, int32_t k, int32_t splits, void* stream) {
    if ((m & 7) || (n & 7)) return ppo_fail(TACO_E_INVALID, "taco_gemm_selftest_mn: m and n must be multiples of 8");
    return gemm_selftest_impl(device, at_bf16, bt_bf16, d_f32, m, n, k, splits, true, stream);
}

}  // extern "C"

// Critic inference behind the C ABI (include/taco_b200.h, taco_critic_*).
//
// Mirrors, for rollout inference only (the PPO update stays in PyTorch), the critic branch of PPO_ActorCritic.act
// (IsaacGymEnvs/algorithms/nets_asymmetry.py:350-352) in the configuration the reference trains with (README.md:60-66:
// --use_critic_encoder=True --critic_encoder_type=LSTM --lenStates=5):
//   LSTMEncoder.forward   nets_asymmetry.py:128-136   nn.LSTM(input, hidden, num_layers, batch_first=True), zero initial state,
//                                                     output = top layer's h after the last time step
//   MLP.forward           nets_asymmetry.py:23-39     [Linear -> ReLU] x L -> Linear -> Identity (:318)
//
// Two kernels compute the same function: an FP32 CUDA-core path (any shape; the parity path) and the tcgen05 bf16 path in
// critic_tc.cuh (one LSTM layer of width <= 64, MLP widths multiples of 64 up to 256).
#include "launch_count.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/taco_b200.h"
#include "critic_tc.cuh"

namespace taco {
int fail(int code, const std::string& msg);      // taco_env.cu
namespace critic {

static int cfail(int code, const std::string& msg) { return taco::fail(code, msg); }

#define CRT_CUDA(expr)                                                                            \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return cfail(TACO_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));        \
    } while (0)

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

constexpr int kMaxLstmLayers = 4;
constexpr int kMaxMlpLayers = 8;     // linear layers of the FP32 path
constexpr int kFpEnvs = 32;          // envs per CTA of the FP32 kernel (one warp lane per env)
constexpr int kFpThreads = 128;
constexpr int kFpTile = 8;           // output neurons per warp pass of the FP32 kernel's MLP part

struct FpParams {
    const float* states; float* value;
    int n_rows, in_dim, seq_len, hidden, lstm_layers, n_mlp;
    const float* w_ih[kMaxLstmLayers]; const float* w_hh[kMaxLstmLayers];
    const float* b_ih[kMaxLstmLayers]; const float* b_hh[kMaxLstmLayers];
    int mlp_sizes[kMaxMlpLayers + 1];
    const float* w[kMaxMlpLayers]; const float* b[kMaxMlpLayers];
    int stride;                       // floats per env in shared memory (odd)
    int off_lstm, off_mlp, mlp_w;     // offsets inside an env's row: per-layer [h_a | h_b | c], then the MLP ping-pong
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// FP32 reference path: CTA = 32 envs (lane = env); everything an env needs sits in its shared-memory row; each warp owns a
// strip of hidden units / output neurons; dot products run over k in order with one FMA per term, the two matrix products
// of a gate are summed like torch (W_ih x + b_ih) + (W_hh h + b_hh).
__global__ void __launch_bounds__(kFpThreads) critic_fp32_kernel(const FpParams p) {
    extern __shared__ float s_row[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kFpThreads / 32;
    const long long row = (long long)blockIdx.x * kFpEnvs + lane;
    const bool valid = row < p.n_rows;
    float* me = s_row + lane * p.stride;
    const int H = p.hidden, T = p.seq_len, In = p.in_dim;
    for (int k = warp; k < T * In; k += kWarps) me[k] = valid ? __ldg(p.states + row * (long long)(T * In) + k) : 0.0f;
    for (int k = warp; k < p.lstm_layers * 3 * H; k += kWarps) me[p.off_lstm + k] = 0.0f;
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        const int cur = t & 1;
        for (int l = 0; l < p.lstm_layers; ++l) {
            float* hl = me + p.off_lstm + l * 3 * H;
            const float* inp = l == 0 ? me + t * In : me + p.off_lstm + (l - 1) * 3 * H + (cur ^ 1) * H;   // layer below, this step
            const int in_l = l == 0 ? In : H;
            const float* hprev = hl + cur * H;
            float* hnew = hl + (cur ^ 1) * H;
            float* cst = hl + 2 * H;
            const float* __restrict__ Wi = p.w_ih[l];
            const float* __restrict__ Wh = p.w_hh[l];
            // two hidden units per pass = 8 gate rows per product; float4 weights along k when the rows are 16-byte aligned.  Every
            // dot product still runs over k in ascending order with one FMA per term.
            const bool vec_i = (in_l & 3) == 0, vec_h = (H & 3) == 0;
            for (int j0 = warp * 2; j0 < H; j0 += kWarps * 2) {
                float a[8], g[8];
                const float* wi[8];
                const float* wh[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int j = min(j0 + (q >> 2), H - 1), row_ = (q & 3) * H + j;
                    a[q] = 0.f; g[q] = 0.f; wi[q] = Wi + (size_t)row_ * in_l; wh[q] = Wh + (size_t)row_ * H;
                }
                if (vec_i) {
                    for (int k = 0; k < in_l; k += 4) {
                        const float x0 = inp[k], x1 = inp[k + 1], x2 = inp[k + 2], x3 = inp[k + 3];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 w = __ldg(reinterpret_cast<const float4*>(wi[q] + k));
                            a[q] = fmaf(x3, w.w, fmaf(x2, w.z, fmaf(x1, w.y, fmaf(x0, w.x, a[q]))));
                        }
                    }
                } else {
                    for (int k = 0; k < in_l; ++k) {
                        const float x = inp[k];
#pragma unroll
                        for (int q = 0; q < 8; ++q) a[q] = fmaf(x, __ldg(wi[q] + k), a[q]);
                    }
                }
                if (vec_h) {
                    for (int k = 0; k < H; k += 4) {
                        const float x0 = hprev[k], x1 = hprev[k + 1], x2 = hprev[k + 2], x3 = hprev[k + 3];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 w = __ldg(reinterpret_cast<const float4*>(wh[q] + k));
                            g[q] = fmaf(x3, w.w, fmaf(x2, w.z, fmaf(x1, w.y, fmaf(x0, w.x, g[q]))));
                        }
                    }
                } else {
                    for (int k = 0; k < H; ++k) {
                        const float x = hprev[k];
#pragma unroll
                        for (int q = 0; q < 8; ++q) g[q] = fmaf(x, __ldg(wh[q] + k), g[q]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int j = j0 + u;
                    if (j < H) {
                        float gate[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) gate[q] = (a[u * 4 + q] + __ldg(p.b_ih[l] + q * H + j)) + (g[u * 4 + q] + __ldg(p.b_hh[l] + q * H + j));
                        const float c = sigmoid_f(gate[1]) * cst[j] + sigmoid_f(gate[0]) * tanhf(gate[2]);
                        cst[j] = c;
                        hnew[j] = sigmoid_f(gate[3]) * tanhf(c);
                    }
                }
            }
            __syncthreads();
        }
    }
    // encoder output: top layer's h after the last step
    float* cur = me + p.off_mlp;
    float* nxt = cur + p.mlp_w;
    {
        const float* hT = me + p.off_lstm + (p.lstm_layers - 1) * 3 * H + (T & 1) * H;
        for (int k = warp; k < H; k += kWarps) cur[k] = hT[k];
    }
    __syncthreads();
    for (int l = 0; l < p.n_mlp; ++l) {
        const int in = p.mlp_sizes[l], out = p.mlp_sizes[l + 1];
        const float* __restrict__ W = p.w[l];
        const float* __restrict__ B = p.b[l];
        const bool last = (l + 1 == p.n_mlp);
        const bool vec = (in & 3) == 0;
        for (int j0 = warp * kFpTile; j0 < out; j0 += kWarps * kFpTile) {
            float acc[kFpTile];
            const float* wr[kFpTile];
#pragma unroll
            for (int i = 0; i < kFpTile; ++i) { acc[i] = 0.f; wr[i] = W + (size_t)min(j0 + i, out - 1) * in; }
            if (vec) {
                for (int k = 0; k < in; k += 4) {
                    const float x0 = cur[k], x1 = cur[k + 1], x2 = cur[k + 2], x3 = cur[k + 3];
#pragma unroll
                    for (int i = 0; i < kFpTile; ++i) {
                        const float4 w = __ldg(reinterpret_cast<const float4*>(wr[i] + k));
                        acc[i] = fmaf(x3, w.w, fmaf(x2, w.z, fmaf(x1, w.y, fmaf(x0, w.x, acc[i]))));
                    }
                }
            } else {
                for (int k = 0; k < in; ++k) {
                    const float x = cur[k];
#pragma unroll
                    for (int i = 0; i < kFpTile; ++i) acc[i] = fmaf(x, __ldg(wr[i] + k), acc[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < kFpTile; ++i) {
                if (j0 + i < out) {
                    const float v = acc[i] + __ldg(B + j0 + i);
                    nxt[j0 + i] = last ? v : fmaxf(v, 0.0f);
                }
            }
        }
        __syncthreads();
        float* tsw = cur; cur = nxt; nxt = tsw;
    }
    if (warp == 0 && valid) p.value[row] = cur[0];
}

}  // namespace critic
}  // namespace taco

using namespace taco::critic;

struct TacoCritic {
    int device = 0;
    int in_dim = 0, seq_len = 0, hidden = 0, lstm_layers = 0;
    std::vector<int> mlp_sizes;          // [hidden, h1, ..., 1]
    int n_mlp = 0;
    std::vector<size_t> lstm_off;        // per layer: float offsets of w_ih, w_hh, b_ih, b_hh (4 per layer)
    std::vector<size_t> w_off, b_off;
    float* lstm_f32 = nullptr;
    float* w_f32 = nullptr;
    float* b_f32 = nullptr;
    bool loaded = false;
    int fp_smem = 0, fp_stride = 0, off_lstm = 0, off_mlp = 0, mlp_w = 0;
    bool tc_ok = false;
    std::string tc_why;
    uint8_t* wimg = nullptr;
    float* bias_pad = nullptr;           // [kMaxHidden][kMaxN]
    float* b_out = nullptr;              // [kOutPad]
    taco::actor::TcLayer tc_layer[2 + kMaxMlpHidden];
    int num_sms = 148;
};

extern "C" {

int taco_critic_destroy(TacoCritic* c);

int taco_critic_create(int device, int32_t in_dim, int32_t seq_len, int32_t lstm_hidden, int32_t lstm_layers, const int32_t* mlp_sizes,
                       int32_t n_mlp_sizes, TacoCritic** out) {
    if (!mlp_sizes || !out) return cfail(TACO_E_INVALID, "taco_critic_create: null argument");
    *out = nullptr;
    if (in_dim < 1 || in_dim > 256 || seq_len < 1 || seq_len > 64) return cfail(TACO_E_INVALID, "taco_critic_create: in_dim must be in [1, 256], seq_len in [1, 64]");
    if (lstm_hidden < 1 || lstm_hidden > 256 || lstm_layers < 1 || lstm_layers > kMaxLstmLayers)
        return cfail(TACO_E_INVALID, "taco_critic_create: lstm_hidden must be in [1, 256], lstm_layers in [1, 4]");
    if (n_mlp_sizes < 2 || n_mlp_sizes > kMaxMlpLayers + 1) return cfail(TACO_E_INVALID, "taco_critic_create: need 2..9 MLP sizes [lstm_hidden, h1, ..., 1]");
    if (mlp_sizes[0] != lstm_hidden) return cfail(TACO_E_INVALID, "taco_critic_create: mlp_sizes[0] must equal lstm_hidden (the encoder output feeds the MLP)");
    if (mlp_sizes[n_mlp_sizes - 1] != 1) return cfail(TACO_E_INVALID, "taco_critic_create: the critic output width is 1");
    for (int i = 0; i < n_mlp_sizes; ++i)
        if (mlp_sizes[i] < 1 || mlp_sizes[i] > 512) return cfail(TACO_E_INVALID, "taco_critic_create: layer sizes must be in [1, 512]");
    int ndev = 0;
    CRT_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return cfail(TACO_E_INVALID, "taco_critic_create: no such CUDA device");
    DevGuard guard(device);
    TacoCritic* c = new (std::nothrow) TacoCritic();
    if (!c) return cfail(TACO_E_NOMEM, "host allocation failed");
    c->device = device; c->in_dim = in_dim; c->seq_len = seq_len; c->hidden = lstm_hidden; c->lstm_layers = lstm_layers;
    c->mlp_sizes.assign(mlp_sizes, mlp_sizes + n_mlp_sizes);
    c->n_mlp = n_mlp_sizes - 1;
    const int H = lstm_hidden;
    size_t lo = 0;
    for (int l = 0; l < lstm_layers; ++l) {
        const int in_l = l == 0 ? in_dim : H;
        const size_t sz[4] = {(size_t)4 * H * in_l, (size_t)4 * H * H, (size_t)4 * H, (size_t)4 * H};
        for (int q = 0; q < 4; ++q) { c->lstm_off.push_back(lo); lo += (sz[q] + 3) & ~(size_t)3; }
    }
    size_t wo = 0, bo = 0;
    int mx = 0;
    for (int l = 0; l < c->n_mlp; ++l) {
        c->w_off.push_back(wo); c->b_off.push_back(bo);
        wo += ((size_t)mlp_sizes[l] * mlp_sizes[l + 1] + 3) & ~(size_t)3;
        bo += (size_t)((mlp_sizes[l + 1] + 3) & ~3);
    }
    for (int i = 0; i < n_mlp_sizes; ++i) mx = mlp_sizes[i] > mx ? mlp_sizes[i] : mx;
    c->off_lstm = seq_len * in_dim;
    c->off_mlp = c->off_lstm + lstm_layers * 3 * H;
    c->mlp_w = mx;
    c->fp_stride = (c->off_mlp + 2 * mx) | 1;
    c->fp_smem = kFpEnvs * c->fp_stride * (int)sizeof(float);
    cudaDeviceProp prop;
    cudaError_t ce = cudaGetDeviceProperties(&prop, device);
    if (ce == cudaSuccess && (size_t)c->fp_smem > prop.sharedMemPerBlockOptin) {
        delete c;
        return cfail(TACO_E_INVALID, "taco_critic_create: shape needs more shared memory per CTA than the device has");
    }
    c->num_sms = prop.multiProcessorCount;
    // ---- is the shape eligible for the tcgen05 path?
    const int n_hidden = c->n_mlp - 1;
    c->tc_ok = true;
    if (prop.major != 10) { c->tc_ok = false; c->tc_why = "device is not sm_100"; }
    else if (lstm_layers != 1) { c->tc_ok = false; c->tc_why = "needs a single LSTM layer"; }
    else if (H % 16 != 0 || H > kTcMaxLstmHidden) { c->tc_ok = false; c->tc_why = "LSTM width must be a multiple of 16, <= 64"; }
    else if (in_dim > kTcMaxIn || (in_dim & 1)) { c->tc_ok = false; c->tc_why = "input width must be even, <= 30"; }
    else if (seq_len > kMaxSeq) { c->tc_ok = false; c->tc_why = "sequence length > 8"; }
    else if (n_hidden < 1 || n_hidden > kMaxMlpHidden) { c->tc_ok = false; c->tc_why = "needs 1..3 MLP hidden layers"; }
    else {
        for (int l = 1; l <= n_hidden; ++l)
            if (mlp_sizes[l] % 64 != 0 || mlp_sizes[l] > taco::actor::kMaxN) { c->tc_ok = false; c->tc_why = "MLP hidden widths must be multiples of 64, <= 256"; }
    }
    if (ce == cudaSuccess) ce = cudaMalloc(&c->lstm_f32, lo * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMalloc(&c->w_f32, wo * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMalloc(&c->b_f32, bo * sizeof(float));
    if (ce == cudaSuccess && c->tc_ok) {
        size_t img = 0;
        // [0] the LSTM step: n = 4H gate rows, K = [h (64) | x (64)] = 2 chunks
        c->tc_layer[0].n = 4 * H; c->tc_layer[0].kchunks = 2; c->tc_layer[0].img_off = 0;
        img += (size_t)2 * 4 * H * 128;
        for (int l = 0; l <= n_hidden; ++l) {          // MLP linear layer l: hidden layers, then the 16-row padded output layer
            const int k = mlp_sizes[l], nn = l < n_hidden ? mlp_sizes[l + 1] : taco::actor::kOutN;
            const int kch = (k + taco::actor::kKC - 1) / taco::actor::kKC;
            c->tc_layer[1 + l].n = nn; c->tc_layer[1 + l].kchunks = kch; c->tc_layer[1 + l].img_off = (uint32_t)img;
            img += (size_t)kch * nn * 128;
        }
        ce = cudaMalloc(&c->wimg, img);
        if (ce == cudaSuccess) ce = cudaMalloc(&c->bias_pad, taco::actor::kMaxHidden * taco::actor::kMaxN * sizeof(float));
        if (ce == cudaSuccess) ce = cudaMalloc(&c->b_out, taco::actor::kOutPad * sizeof(float));
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(critic_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCriticSmemBytes);
    }
    if (ce == cudaSuccess && c->fp_smem > 48 * 1024)
        ce = cudaFuncSetAttribute(critic_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, c->fp_smem);
    if (ce != cudaSuccess) {
        const std::string msg = std::string("taco_critic_create: ") + cudaGetErrorString(ce);
        taco_critic_destroy(c);
        return cfail(ce == cudaErrorMemoryAllocation ? TACO_E_NOMEM : TACO_E_CUDA, msg);
    }
    *out = c;
    return TACO_OK;
}

int taco_critic_destroy(TacoCritic* c) {
    if (!c) return TACO_OK;
    DevGuard guard(c->device);
    cudaFree(c->lstm_f32); cudaFree(c->w_f32); cudaFree(c->b_f32);
    cudaFree(c->wimg); cudaFree(c->bias_pad); cudaFree(c->b_out);
    delete c;
    return TACO_OK;
}

int taco_critic_load(TacoCritic* c, const float* const* lstm_host, const float* const* mlp_weights_host, const float* const* mlp_biases_host,
                     void* stream) {
    if (!c || !lstm_host || !mlp_weights_host || !mlp_biases_host) return cfail(TACO_E_INVALID, "taco_critic_load: null argument");
    for (int i = 0; i < 4 * c->lstm_layers; ++i)
        if (!lstm_host[i]) return cfail(TACO_E_INVALID, "taco_critic_load: null LSTM parameter pointer");
    for (int l = 0; l < c->n_mlp; ++l)
        if (!mlp_weights_host[l] || !mlp_biases_host[l]) return cfail(TACO_E_INVALID, "taco_critic_load: null MLP layer pointer");
    DevGuard guard(c->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int H = c->hidden;
    for (int l = 0; l < c->lstm_layers; ++l) {
        const int in_l = l == 0 ? c->in_dim : H;
        const size_t sz[4] = {(size_t)4 * H * in_l, (size_t)4 * H * H, (size_t)4 * H, (size_t)4 * H};
        for (int q = 0; q < 4; ++q)
            CRT_CUDA(cudaMemcpyAsync(c->lstm_f32 + c->lstm_off[4 * l + q], lstm_host[4 * l + q], sz[q] * sizeof(float), cudaMemcpyDefault, s));
    }
    for (int l = 0; l < c->n_mlp; ++l) {
        const int in = c->mlp_sizes[l], out = c->mlp_sizes[l + 1];
        CRT_CUDA(cudaMemcpyAsync(c->w_f32 + c->w_off[l], mlp_weights_host[l], (size_t)in * out * sizeof(float), cudaMemcpyDefault, s));
        CRT_CUDA(cudaMemcpyAsync(c->b_f32 + c->b_off[l], mlp_biases_host[l], (size_t)out * sizeof(float), cudaMemcpyDefault, s));
    }
    if (c->tc_ok) {
        const int n_hidden = c->n_mlp - 1;
        CRT_CUDA(cudaMemsetAsync(c->bias_pad, 0, taco::actor::kMaxHidden * taco::actor::kMaxN * sizeof(float), s));
        pack_lstm_kernel<<<64, 256, 0, s>>>(c->lstm_f32 + c->lstm_off[0], c->lstm_f32 + c->lstm_off[1], c->lstm_f32 + c->lstm_off[2],
                                            c->lstm_f32 + c->lstm_off[3], H, c->in_dim, 4 * H, c->wimg); TACO_LAUNCHED();
        for (int l = 0; l <= n_hidden; ++l) {
            const int in = c->mlp_sizes[l], out = c->mlp_sizes[l + 1];
            taco::actor::pack_weights_kernel<<<64, 256, 0, s>>>(c->w_f32 + c->w_off[l], out, c->tc_layer[1 + l].n, in, c->wimg + c->tc_layer[1 + l].img_off); TACO_LAUNCHED();
            if (l < n_hidden)
                CRT_CUDA(cudaMemcpyAsync(c->bias_pad + (size_t)(1 + l) * taco::actor::kMaxN, c->b_f32 + c->b_off[l], (size_t)out * sizeof(float),
                                         cudaMemcpyDeviceToDevice, s));
        }
        CRT_CUDA(cudaMemsetAsync(c->b_out, 0, taco::actor::kOutPad * sizeof(float), s));
        CRT_CUDA(cudaMemcpyAsync(c->b_out, c->b_f32 + c->b_off[n_hidden], sizeof(float), cudaMemcpyDeviceToDevice, s));
        CRT_CUDA(cudaGetLastError());
    }
    CRT_CUDA(cudaStreamSynchronize(s));     // the host buffers may be released by the caller on return
    c->loaded = true;
    return TACO_OK;
}

int taco_critic_tc_available(TacoCritic* c) { return (c && c->tc_ok) ? 1 : 0; }

int taco_critic_forward(TacoCritic* c, const float* states_dev, float* value_dev, int32_t n, int32_t use_tensor_cores, void* stream) {
    if (!c || !states_dev || !value_dev) return cfail(TACO_E_INVALID, "taco_critic_forward: null argument");
    if (!c->loaded) return cfail(TACO_E_INVALID, "taco_critic_forward: call taco_critic_load first");
    if (n <= 0) return cfail(TACO_E_INVALID, "taco_critic_forward: n must be positive");
    DevGuard guard(c->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (use_tensor_cores) {
        if (!c->tc_ok) return cfail(TACO_E_INVALID, "taco_critic_forward: tensor-core path unavailable for this shape: " + c->tc_why);
        CriticTcParams p;
        memset(&p, 0, sizeof(p));
        p.states = states_dev; p.value = value_dev;
        p.in_dim = c->in_dim; p.seq_len = c->seq_len; p.n_rows = n; p.num_tiles = (n + taco::actor::kTileM - 1) / taco::actor::kTileM;
        p.lstm_hidden = c->hidden; p.n_hidden = c->n_mlp - 1;
        p.wimg = c->wimg; p.bias = c->bias_pad; p.b_out = c->b_out;
        for (int l = 0; l <= p.n_hidden + 1; ++l) p.layer[l] = c->tc_layer[l];
        p.tiles_per_cta = p.num_tiles <= c->num_sms ? 1 : 2;
        static const int stagger = [] { const char* e = getenv("TACO_CRITIC_STAGGER"); return e ? atoi(e) : taco::critic::kCriticStagger; }();
        p.stagger = stagger;
        const int num_pairs = (p.num_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
        const int grid = num_pairs < c->num_sms ? num_pairs : c->num_sms;
        // developer aid: TACO_CRITIC_TIMELINE=<file> records clock64 stamps of CTA 0's MMA issuer / epilogue (synchronous)
        const char* tl = getenv("TACO_CRITIC_TIMELINE");
        if (tl && *tl) {
            CRT_CUDA(cudaMalloc(&p.dbg, 3 * taco::actor::kDbgCap * sizeof(unsigned long long)));
            CRT_CUDA(cudaMemsetAsync(p.dbg, 0, 3 * taco::actor::kDbgCap * sizeof(unsigned long long), s));
        }
        critic_tc_kernel<<<grid, taco::actor::kTcThreads, kCriticSmemBytes, s>>>(p); TACO_LAUNCHED();
        if (p.dbg) {
            std::vector<unsigned long long> h(3 * taco::actor::kDbgCap);
            CRT_CUDA(cudaStreamSynchronize(s));
            CRT_CUDA(cudaMemcpy(h.data(), p.dbg, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            cudaFree(p.dbg);
            if (FILE* f = fopen(tl, "wb")) { fwrite(h.data(), sizeof(unsigned long long), h.size(), f); fclose(f); }
        }
    } else {
        FpParams p;
        memset(&p, 0, sizeof(p));
        p.states = states_dev; p.value = value_dev; p.n_rows = n;
        p.in_dim = c->in_dim; p.seq_len = c->seq_len; p.hidden = c->hidden; p.lstm_layers = c->lstm_layers; p.n_mlp = c->n_mlp;
        for (int l = 0; l < c->lstm_layers; ++l) {
            p.w_ih[l] = c->lstm_f32 + c->lstm_off[4 * l + 0]; p.w_hh[l] = c->lstm_f32 + c->lstm_off[4 * l + 1];
            p.b_ih[l] = c->lstm_f32 + c->lstm_off[4 * l + 2]; p.b_hh[l] = c->lstm_f32 + c->lstm_off[4 * l + 3];
        }
        for (int l = 0; l <= c->n_mlp; ++l) p.mlp_sizes[l] = c->mlp_sizes[l];
        for (int l = 0; l < c->n_mlp; ++l) { p.w[l] = c->w_f32 + c->w_off[l]; p.b[l] = c->b_f32 + c->b_off[l]; }
        p.stride = c->fp_stride; p.off_lstm = c->off_lstm; p.off_mlp = c->off_mlp; p.mlp_w = c->mlp_w;
        critic_fp32_kernel<<<(n + kFpEnvs - 1) / kFpEnvs, kFpThreads, c->fp_smem, s>>>(p); TACO_LAUNCHED();
    }
    CRT_CUDA(cudaGetLastError());
    return TACO_OK;
}

}  // extern "C"

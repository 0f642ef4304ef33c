// probe: throughput of MUFU.TANH (tanh.approx.f32, tanh.approx.f16x2) and MUFU.EX2 per SM, next to FFMA
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define ITERS 2048
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a) {
    float x[8];
    for (int j = 0; j < 8; ++j) x[j] = threadIdx.x * 1e-3f + j * 0.1f;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[j])); }
            else if (MODE == 1) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j])); }
            else if (MODE == 2) { unsigned u = __float_as_uint(x[j]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); x[j] = __uint_as_float(u); }
            else { x[j] = fmaf(x[j], a, 0.25f); }
        }
    }
    float s = 0; for (int j = 0; j < 8; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 0.999f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, 0.999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 8 * 256 * ITERS * 8;
    printf("%-24s %.3f ms  %.1f thread-instr/clk/SM @1.9GHz\n", name, ms, ops / (ms * 1e-3) / 148 / 1.9e9);
}
int main() { run<0>("tanh.approx.f32"); run<1>("ex2.approx.f32"); run<2>("tanh.approx.f16x2"); run<3>("ffma"); return 0; }

// probe: issue cost of packed f32x2 arithmetic (FMUL2/FADD2/FFMA2) vs scalar on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    unsigned long long r; float2 z = make_float2(-0.0f, -0.0f); asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&z)));
    return *reinterpret_cast<float2*>(&r); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    unsigned long long r; float2 o = make_float2(1.0f, 1.0f); asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&o)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return *reinterpret_cast<float2*>(&r); }
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, float a, float b) {
    float x[8];
    for (int j = 0; j < 8; ++j) x[j] = threadIdx.x * 1e-3f + j;
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {        // scalar mul+add, 8 chains: 16 instr / 16 flop-ops
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = __fadd_rn(__fmul_rn(x[j], a), b);
        } else if (MODE == 1) { // packed mul2+add2, 4 pair chains: 8 instr / 16 ops
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                float2 v = make_float2(x[j], x[j + 1]);
                v = add2(mul2(v, make_float2(a, a)), make_float2(b, b));
                x[j] = v.x; x[j + 1] = v.y;
            }
        } else if (MODE == 2) { // scalar fma 8 chains
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = __fmaf_rn(x[j], a, b);
        } else {                // packed fma2 4 pair chains
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                float2 v = make_float2(x[j], x[j + 1]);
                v = __ffma2_rn(v, make_float2(a, a), make_float2(b, b));
                x[j] = v.x; x[j + 1] = v.y;
            }
        }
    }
    float s = 0; for (int j = 0; j < 8; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, int ops_per_iter) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 0.999f, 1e-3f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, 0.999f, 1e-3f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double elem_ops = 148.0 * 8 * 256 * ITERS * ops_per_iter;
    printf("%-28s %.3f ms  %.2f Gop/s element ops (mul or add or fma counted as 1)  -> %.1f elem-ops/clk/SM @1.9GHz\n", name, ms, elem_ops / ms * 1e-6, elem_ops / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
    run<0>("scalar FMUL+FADD", 16);
    run<1>("packed FMUL2+FADD2", 16);
    run<2>("scalar FFMA", 8);
    run<3>("packed FFMA2", 8);
    return 0;
}

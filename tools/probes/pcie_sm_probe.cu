// probe: bandwidth of SM loads / stores that target pinned HOST memory (zero-copy over PCIe), by access width,
// next to the copy engine.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pcie_sm_probe pcie_sm_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <typename T> __global__ void wr(T* dst, size_t n, T v) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}
__global__ void rd16(const float4* src, size_t n, float* sink) {
    float acc = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) { float4 v = __ldg(src + i); acc += v.x + v.y + v.z + v.w; }
    if (acc == 1234.5f) *sink = acc;
}
// the env step's I/O shape: per thread 16 B in, 4 + 8 + 1 B out
__global__ void mix(const float4* src, float* o4, long long* o8, uint8_t* o1, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = __ldg(src + i); o4[i] = v.x; o8[i] = v.y > 0.f; o1[i] = v.z > 0.f;
    }
}
// same, but the 1-byte flags of 4 neighbouring threads are merged by shuffles into one 32-bit store
__global__ void mix_packed(const float4* src, float* o4, long long* o8, uint32_t* o1, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = __ldg(src + i); o4[i] = v.x; o8[i] = v.y > 0.f;
        uint32_t b = v.z > 0.f;
        b |= __shfl_down_sync(0xffffffffu, b, 1) << 8; b |= __shfl_down_sync(0xffffffffu, b, 2) << 16;
        if ((threadIdx.x & 3) == 0) o1[i >> 2] = b;
    }
}
template <typename F> float timed(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); for (int r = 0; r < 5; ++r) f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
int main() {
    const size_t n = 2097152;       // "envs"
    void *h_in, *h4, *h8, *h1; float* sink; void* dbuf;
    cudaHostAlloc(&h_in, n * 16, cudaHostAllocMapped); cudaHostAlloc(&h4, n * 16, cudaHostAllocMapped);
    cudaHostAlloc(&h8, n * 16, cudaHostAllocMapped); cudaHostAlloc(&h1, n * 16, cudaHostAllocMapped);
    cudaMalloc(&sink, 4); cudaMalloc(&dbuf, n * 16);
    const int g = 148 * 8, b = 256;
    float ms;
    ms = timed([&] { cudaMemcpyAsync(dbuf, h_in, n * 16, cudaMemcpyHostToDevice); });  printf("copy engine H2D 32 MiB      %.3f ms %.1f GB/s\n", ms, n * 16 / ms * 1e-6);
    ms = timed([&] { cudaMemcpyAsync(h4, dbuf, n * 16, cudaMemcpyDeviceToHost); });    printf("copy engine D2H 32 MiB      %.3f ms %.1f GB/s\n", ms, n * 16 / ms * 1e-6);
    ms = timed([&] { rd16<<<g, b>>>((const float4*)h_in, n, sink); });                 printf("SM read 16 B/thread         %.3f ms %.1f GB/s\n", ms, n * 16 / ms * 1e-6);
    ms = timed([&] { wr<uint8_t><<<g, b>>>((uint8_t*)h1, n * 4, 1); });                printf("SM write 1 B/thread (8 MiB) %.3f ms %.1f GB/s\n", ms, n * 4 / ms * 1e-6);
    ms = timed([&] { wr<float><<<g, b>>>((float*)h4, n * 4, 1.f); });                  printf("SM write 4 B/thread         %.3f ms %.1f GB/s\n", ms, n * 16 / ms * 1e-6);
    ms = timed([&] { wr<long long><<<g, b>>>((long long*)h8, n * 2, 1ll); });          printf("SM write 8 B/thread         %.3f ms %.1f GB/s\n", ms, n * 16 / ms * 1e-6);
    ms = timed([&] { wr<float4><<<g, b>>>((float4*)h8, n, make_float4(1, 2, 3, 4)); }); printf("SM write 16 B/thread        %.3f ms %.1f GB/s\n", ms, n * 16 / ms * 1e-6);
    ms = timed([&] { mix<<<g, b>>>((const float4*)h_in, (float*)h4, (long long*)h8, (uint8_t*)h1, n); });
    printf("SM mix 16 in + 4+8+1 out    %.3f ms  in %.1f GB/s out %.1f GB/s\n", ms, n * 16 / ms * 1e-6, n * 13 / ms * 1e-6);
    ms = timed([&] { mix_packed<<<g, b>>>((const float4*)h_in, (float*)h4, (long long*)h8, (uint32_t*)h1, n); });
    printf("SM mix, flags packed x4     %.3f ms  in %.1f GB/s out %.1f GB/s\n", ms, n * 16 / ms * 1e-6, n * 13 / ms * 1e-6);
    ms = timed([&] { mix<<<g, b>>>((const float4*)h_in, (float*)h4, (long long*)h8, (uint8_t*)dbuf, n); });
    printf("SM mix, flags to device     %.3f ms  in %.1f GB/s out %.1f GB/s\n", ms, n * 16 / ms * 1e-6, n * 12 / ms * 1e-6);
    ms = timed([&] { mix<<<g, b>>>((const float4*)h_in, (float*)h4, (long long*)dbuf, (uint8_t*)dbuf + n * 8, n); });
    printf("SM mix, only rew to host    %.3f ms  in %.1f GB/s out %.1f GB/s\n", ms, n * 16 / ms * 1e-6, n * 4 / ms * 1e-6);
    // concurrent copy-engine D2H while SM reads host memory
    cudaStream_t s2; cudaStreamCreate(&s2);
    ms = timed([&] { rd16<<<g, b>>>((const float4*)h_in, n, sink); cudaMemcpyAsync(h8, dbuf, n * 13, cudaMemcpyDeviceToHost, s2); cudaStreamSynchronize(s2); });
    printf("SM read 32 MiB || CE D2H 26 MiB  %.3f ms\n", ms);
    return 0;
}

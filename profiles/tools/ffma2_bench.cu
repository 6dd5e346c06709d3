// ffma2_bench: fp32 FMA throughput of one B200 with scalar FFMA and with packed fma.rn.f32x2
// (FFMA2), 8 independent accumulator chains per thread, 1024 threads per SM resident.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <bool PACKED>
__global__ void __launch_bounds__(256) k(float *out, const float *in, int n) {
    const float a = in[threadIdx.x & 31], b = in[(threadIdx.x & 31) + 32];
    float r = 0.f;
    if (PACKED) {
        unsigned long long acc[8], A, B;
        asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a), "f"(b));
        asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b), "f"(a));
        for (int j = 0; j < 8; ++j) acc[j] = 0;
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j]) : "l"(A), "l"(B));
        for (int j = 0; j < 8; ++j) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[j])); r += x + y; }
    } else {
        float acc[16];
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < 16; ++j) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc[j]) : "f"(a), "f"(b));
        for (int j = 0; j < 16; ++j) r += acc[j];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float *in, *out; cudaMalloc(&in, 256); cudaMalloc(&out, 148 * 8 * 256 * 4); cudaMemset(in, 0, 256);
    const int n = 20000, grid = 148 * 4;
    cudaEvent_t t0, t1; cudaEventCreate(&t0); cudaEventCreate(&t1);
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        cudaEventRecord(t0); k<false><<<grid, 256>>>(out, in, n); cudaEventRecord(t1); cudaEventSynchronize(t1); cudaEventElapsedTime(&ms, t0, t1);
        printf("FFMA : %.3f ms, %.1f TFLOP/s\n", ms, 2.0 * 16 * n * grid * 256 / ms / 1e9);
        cudaEventRecord(t0); k<true><<<grid, 256>>>(out, in, n); cudaEventRecord(t1); cudaEventSynchronize(t1); cudaEventElapsedTime(&ms, t0, t1);
        printf("FFMA2: %.3f ms, %.1f TFLOP/s\n", ms, 2.0 * 16 * n * grid * 256 / ms / 1e9);
    }
    return 0;
}

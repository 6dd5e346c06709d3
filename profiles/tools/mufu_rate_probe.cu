// Special-function-unit throughput probe (roofline of the policy kernel, csrc/b200policy.cu):
// warp instructions per clock per SM of tanh.approx.f32 / tanh.approx.bf16x2 / ex2.approx / rcp.approx,
// and of the cvt.rn.bf16x2.f32 pack, from 16 independent chains per thread, 1024 threads per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/mufu_rate_probe mufu_rate_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024) probe(float *out, int iters) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.001f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 1) { unsigned u = __float_as_uint(v[i]); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
            if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 3) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 4) { unsigned u; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(u) : "f"(v[i])); v[i] = __uint_as_float(u); }
            if (OP == 5) { unsigned u = __float_as_uint(v[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u)); v[i] = __uint_as_float(u); }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char *name, int sms, float mhz) {
    float *out;
    cudaMalloc(&out, sizeof(float) * sms * 2 * 1024);
    const int iters = 4096;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    probe<OP><<<sms * 2, 1024>>>(out, iters);
    cudaEventRecord(a);
    probe<OP><<<sms * 2, 1024>>>(out, iters);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)sms * 2 * 1024 * 16 * iters;
    printf("%-22s %8.3f ms  %7.1f G lane-ops/s  %5.2f lane-ops per clock per SM at %.0f MHz\n", name, ms, ops / ms / 1e6,
           ops / (ms * 1e-3) / sms / (mhz * 1e6), mhz);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const float mhz = khz / 1000.f;
    printf("%s, %d SMs, %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
    run<0>("tanh.approx.f32", p.multiProcessorCount, mhz);
    run<1>("tanh.approx.bf16x2", p.multiProcessorCount, mhz);
    run<5>("tanh.approx.f16x2", p.multiProcessorCount, mhz);
    run<2>("ex2.approx.ftz.f32", p.multiProcessorCount, mhz);
    run<3>("rcp.approx.ftz.f32", p.multiProcessorCount, mhz);
    run<4>("cvt.rn.bf16x2.f32", p.multiProcessorCount, mhz);
    return 0;
}

// How fast does one SM's copy engine retire bulk tensor copies?  (probe, not shipped)
// Every CTA issues N loads of a {32 x 32 x 1} fp32 box (4 KB, swizzle 128B_ATOM_32B) from a
// [E][784][64] tensor into a ring of shared-memory buffers, DEPTH boxes in flight, and reports
// cycles per box; then the same with bulk tensor stores, and with {32 x 1} boxes (128 B).
//   tma_rate_probe <ctas> <depth>
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_rate_probe tma_rate_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int RING = 16;

// mode 0: 3-D loads of 4 KB boxes; mode 1: 3-D stores of 4 KB boxes; mode 2: 2-D loads of 128-byte boxes
__global__ void __launch_bounds__(32) probe(const __grid_constant__ CUtensorMap map3, const __grid_constant__ CUtensorMap map2,
                                            int mode, int n, int depth, int envs, long long *out, int box_bytes, int rows) {
    extern __shared__ __align__(1024) unsigned char raw[];
    __shared__ __align__(8) uint64_t bar[RING];
    unsigned char *sm = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
    if (threadIdx.x == 0) {
        for (int i = 0; i < RING; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int e0 = (blockIdx.x * 37) % envs;
    const long long t0 = clock64();
    if (mode == 0 || mode == 2 || mode == 5) {
        const uint32_t bytes = (uint32_t)box_bytes;
        for (int i = 0; i < n + depth; ++i) {
            if (i >= depth) mbar_wait(&bar[(i - depth) % RING], ((i - depth) / RING) & 1);
            if (i < n) {
                const int b = i % RING;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[b])), "r"(bytes) : "memory");
                const int per_env = 784 / rows;
                const int e = (e0 + i / (2 * per_env)) % envs, f = (i % per_env) * rows, j = ((i / per_env) & 1) * 32;
                if (mode == 0)
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 ::"r"(smem_u32(sm + (b * box_bytes) % (RING * 4096))), "l"(&map3), "r"(j), "r"(f), "r"(e), "r"(smem_u32(&bar[b])) : "memory");
                else if (mode == 5)
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                                 ::"r"(smem_u32(sm + (b * box_bytes) % (RING * 4096))), "l"(&map3), "r"(0), "r"(f), "r"(0), "r"(e), "r"(smem_u32(&bar[b])) : "memory");
                else
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                                 ::"r"(smem_u32(sm + b * 4096)), "l"(&map2), "r"(j), "r"((e * 784 + f + i) % (envs * 784)), "r"(smem_u32(&bar[b])) : "memory");
            }
        }
    } else {
        for (int i = 0; i < n; ++i) {
            const int per_env = 784 / rows;
            const int e = (e0 + i / (2 * per_env)) % envs, f = (i % per_env) * rows, j = ((i / per_env) & 1) * 32;
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                         ::"l"(&map3), "r"(j), "r"(f), "r"(e), "r"(smem_u32(sm + ((i % RING) * box_bytes) % (RING * 4096))) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            // at most `depth` stores in flight (depth is 1, 2, 4 or 8 here)
            if (depth == 1) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            else if (depth == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else if (depth == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    out[blockIdx.x] = clock64() - t0;
}

int main(int argc, char **argv) {
    const int ctas = argc > 1 ? atoi(argv[1]) : 1, depth = argc > 2 ? atoi(argv[2]) : 4, envs = 2048, n = 960;
    float *w;
    const size_t pp = 50892;
    CK(cudaMalloc(&w, envs * pp * 4));
    CK(cudaMemset(w, 0, envs * pp * 4));
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap m3, m2;
    {
        const cuuint64_t dims[3] = {64, 784, (cuuint64_t)envs}, strides[2] = {256, pp * 4};
        const cuuint32_t box[3] = {32, 32, 1}, es[3] = {1, 1, 1};
        if (((encode_fn)fn)(&m3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode3 failed\n"); return 3; }
        const cuuint64_t d2[2] = {64, (cuuint64_t)envs * 784}, s2[1] = {256};
        const cuuint32_t b2[2] = {32, 1}, e2[2] = {1, 1};
        if (((encode_fn)fn)(&m2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, w, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode2 failed\n"); return 3; }
    }
    long long *out, host[2048];
    CK(cudaMalloc(&out, sizeof(host)));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, RING * 4096 + 1024));
    struct Cfg { const char *name; int mode, rows, rank4; };
    const Cfg cfgs[] = {{"load  {32,32,1} 4 KB", 0, 32, 0}, {"load  {32,64,1} 8 KB", 0, 64, 0}, {"load  {32,128,1} 16 KB", 0, 128, 0},
                        {"load  {32,32,2,1} 8 KB (4-d view, both column halves)", 5, 32, 1}, {"load  {32,1} 128 B (2-d)", 2, 1, 0},
                        {"store {32,32,1} 4 KB", 1, 32, 0}, {"store {32,128,1} 16 KB", 1, 128, 0}};
    for (const Cfg &c : cfgs) {
        CUtensorMap m;
        const cuuint32_t es[4] = {1, 1, 1, 1};
        if (c.rank4) {
            const cuuint64_t dims[4] = {32, 784, 2, (cuuint64_t)envs}, strides[3] = {256, 128, pp * 4};
            const cuuint32_t box[4] = {32, (cuuint32_t)c.rows, 2, 1};
            if (((encode_fn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode4 failed\n"); continue; }
        } else {
            const cuuint64_t dims[3] = {64, 784, (cuuint64_t)envs}, strides[2] = {256, pp * 4};
            const cuuint32_t box[3] = {32, (cuuint32_t)c.rows, 1};
            if (((encode_fn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode3 failed\n"); continue; }
        }
        const int box_bytes = c.mode == 2 ? 128 : 128 * c.rows * (c.rank4 ? 2 : 1);
        const int count = c.rows >= 64 ? n / 2 : n;
        for (int rep = 0; rep < 2; ++rep) {
            probe<<<ctas, 32, RING * 4096 + 1024>>>(m, m2, c.mode, count, depth, envs, out, box_bytes, c.mode == 2 ? 32 : c.rows);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(host, out, ctas * sizeof(long long), cudaMemcpyDeviceToHost));
            long long worst = 0, sum = 0;
            for (int i = 0; i < ctas; ++i) { sum += host[i]; if (host[i] > worst) worst = host[i]; }
            if (rep) printf("%-58s ctas %3d depth %d: %7.1f cycles per box, %5.1f bytes/cycle (worst CTA %.1f cycles)\n", c.name, ctas, depth,
                            (double)sum / ctas / count, box_bytes * (double)count * ctas / sum, (double)worst / count);
        }
    }
    return 0;
}

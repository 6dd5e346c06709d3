// Descriptor-convention probe for tcgen05.mma kind::tf32 with MN-major operands (not a test, not
// shipped): one CTA fills shared memory with two known matrices in a candidate canonical layout,
// issues the MMAs of one K = 32 tile and dumps the TMEM accumulator; the host compares it with the
// plain product.  One variant per process (a bad descriptor may fault):
//   tc2_desc_probe <case> <mn_layout_type> <swizzle> <lbo_is_block> <group_rows>
//   case 0: D[128 x 64]  = A (MN-major, M = 128, 4 blocks of 32) . B (K-major SW128, N = 64)   "forward"
//   case 1: D[128 x 128] = A (MN-major) . B (MN-major, N = 128, 4 blocks of 32)                 "backward"
//   mn_layout_type: descriptor bits 61..63 for the MN-major operands (1 = 128B_BASE32B, 2 = 128B, 0 = none)
//   swizzle: how the filler permutes chunks in a 128-byte row: 0 none, 1 = 32 B chunks ^ (row & 3),
//            2 = 16 B chunks ^ (row & 7)
//   lbo_is_block: 1 -> LBO = stride between 32-element MN blocks, SBO = stride between K row groups;
//                 0 -> the other way round
//   group_rows: K rows per group (4 or 8) -> group stride = group_rows * 128 bytes
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc2_desc_probe tc2_desc_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

struct Params {
    int kase, mn_type, swz, lbo_is_block, group_rows;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)type << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t chunk_off(int swz, int row, int col /* element 0..31 */) {
    const int byte = col * 4;
    if (swz == 1) return (uint32_t)(((((byte >> 5) ^ (row & 3)) << 5) | (byte & 31)));
    if (swz == 2) return (uint32_t)(((((byte >> 4) ^ (row & 7)) << 4) | (byte & 15)));
    return (uint32_t)byte;
}

// logical inputs: A[m][k] (128 x 32), B[n][k] (N x 32); output D[m][n]
__global__ void __launch_bounds__(128) probe(const float *A, const float *Bm, float *D, Params p) {
    extern __shared__ __align__(1024) unsigned char raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    unsigned char *sm = raw + ((1024 - (smem_u32(raw) & 1023)) & 1023);
    unsigned char *sa = sm, *sb = sm + 16384;
    const int N = p.kase == 0 ? 64 : 128;
    // A: MN-major blocks [K rows = 32][128 B = 32 m], block b = m / 32 at b * 4096
    for (int i = tid; i < 128 * 32; i += 128) {
        const int m = i >> 5, k = i & 31;
        *reinterpret_cast<float *>(sa + (m >> 5) * 4096 + k * 128 + chunk_off(p.swz, k, m & 31)) = A[m * 32 + k];
    }
    if (p.kase == 0) {       // B: K-major SW128: rows n, 128 B = 32 k, 16 B chunks ^ (n & 7)
        for (int i = tid; i < N * 32; i += 128) {
            const int n = i >> 5, k = i & 31;
            *reinterpret_cast<float *>(sb + n * 128 + chunk_off(2, n, k)) = Bm[n * 32 + k];
        }
    } else {                 // B: MN-major like A
        for (int i = tid; i < N * 32; i += 128) {
            const int n = i >> 5, k = i & 31;
            *reinterpret_cast<float *>(sb + (n >> 5) * 4096 + k * 128 + chunk_off(p.swz, k, n & 31)) = Bm[n * 32 + k];
        }
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t grp = (uint32_t)p.group_rows * 128, blk = 4096;
        const uint32_t lbo = p.lbo_is_block ? blk : grp, sbo = p.lbo_is_block ? grp : blk;
        uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        if (p.kase == 1) idesc |= 1u << 16;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = make_desc(smem_u32(sa) + kk * 1024, lbo, sbo, (uint32_t)p.mn_type);
            const uint64_t bd = p.kase == 0 ? make_desc(smem_u32(sb) + kk * 32, 16, 1024, 2u)
                                            : make_desc(smem_u32(sb) + kk * 1024, lbo, sbo, (uint32_t)p.mn_type);
            mma_tf32(tmem, ad, bd, idesc, kk ? 1u : 0u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t done = 0;
        long long t0 = clock64();
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
            if (clock64() - t0 > 2000000000LL) { if (tid == 0) printf("timeout waiting for the MMAs\n"); break; }
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) D[tid * N + c0 + i] = __uint_as_float(v[i]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main(int argc, char **argv) {
    if (argc < 6) { printf("usage: %s case mn_type swz lbo_is_block group_rows\n", argv[0]); return 1; }
    Params p = {atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5])};
    const int N = p.kase == 0 ? 64 : 128;
    float *hA = (float *)malloc(128 * 32 * 4), *hB = (float *)malloc(N * 32 * 4), *hD = (float *)malloc(128 * N * 4);
    srand(7);
    auto rnd = []() {           // tf32-exact values
        float v = (float)(rand() % 2001 - 1000) / 256.0f;
        uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; memcpy(&v, &u, 4); return v;
    };
    for (int i = 0; i < 128 * 32; ++i) hA[i] = rnd();
    for (int i = 0; i < N * 32; ++i) hB[i] = rnd();
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, 128 * 32 * 4)); CK(cudaMalloc(&dB, N * 32 * 4)); CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMemcpy(dA, hA, 128 * 32 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB, N * 32 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, 128 * N * 4));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 34816));
    probe<<<1, 128, 34816>>>(dA, dB, dD, p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hD, dD, 128 * N * 4, cudaMemcpyDeviceToHost));
    double worst = 0.0, scale = 0.0;
    int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double ref = 0.0;
            for (int k = 0; k < 32; ++k) ref += (double)hA[m * 32 + k] * (double)hB[n * 32 + k];
            const double err = fabs(ref - (double)hD[m * N + n]);
            if (err > worst) worst = err;
            if (fabs(ref) > scale) scale = fabs(ref);
            if (err > 1e-3 * (1.0 + fabs(ref))) ++bad;
        }
    printf("case %d mn_type %d swz %d lbo_is_block %d group_rows %d : max err %.3e (max |ref| %.1f), %d / %d entries wrong -> %s\n",
           p.kase, p.mn_type, p.swz, p.lbo_is_block, p.group_rows, worst, scale, bad, 128 * N, bad ? "MISMATCH" : "MATCH");
    return 0;
}

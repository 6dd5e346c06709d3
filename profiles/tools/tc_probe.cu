// tc_probe: characterises tcgen05.mma kind::tf32 on sm_100a before it is used in the eval
// kernel.  One CTA, operands written by the threads into the no-swizzle canonical layouts
// (K-major and MN-major), accumulator read back from TMEM with tcgen05.ld.  Prints
//   * whether D matches A.B^T for each (M, N, A-major, B-major) under the assumed layouts,
//   * the TMEM lane every accumulator row lands in (M = 64 vs 128),
//   * whether fp32 inputs are truncated or rounded to tf32,
//   * the error of the 3xTF32 split against an fp64 reference.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                     // descriptor version (sm_100)
    return d;                                   // layout_type 0 = no swizzle, base_offset 0
}

// byte offset of element (r, k) of an R x K operand
//   K-major : core matrix = 8 rows x 16 bytes; [r/8][k/4][r%8][k%4]   -> LBO = 128, SBO = (K/4)*128
//   MN-major: core matrix = 8 k    x 16 bytes; [k/8][r/4][k%8][r%4]   -> SBO = 128, LBO = (R/4)*128
__host__ __device__ inline int off_kmajor(int r, int k, int K) { return ((r / 8) * (K / 4) + k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4; }
__host__ __device__ inline int off_mnmajor(int r, int k, int R) { return ((k / 8) * (R / 4) + r / 4) * 128 + (k % 8) * 16 + (r % 4) * 4; }

struct Params {
    int M, N, K, a_mn, b_mn, nsplit;            // nsplit 1: plain tf32, 3: 3xTF32
    const float *A, *B;                         // logical row-major A[M][K], B[N][K]
    float *D;                                   // [128 lanes][N] as read from TMEM
};

__global__ void __launch_bounds__(128) probe(Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t mbar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int abytes = p.M * p.K * 4, bbytes = p.N * p.K * 4;
    unsigned char *Ahi = smem, *Alo = smem + abytes, *Bhi = smem + 2 * abytes, *Blo = Bhi + bbytes;
    for (int i = tid; i < p.M * p.K; i += 128) {
        const int m = i / p.K, k = i % p.K;
        const float v = p.A[i];
        const float hi = p.nsplit == 3 ? __uint_as_float(__float_as_uint(v) & 0xFFFFE000u) : v;
        const int o = p.a_mn ? off_mnmajor(m, k, p.M) : off_kmajor(m, k, p.K);
        *(float *)(Ahi + o) = hi;
        *(float *)(Alo + o) = v - hi;
    }
    for (int i = tid; i < p.N * p.K; i += 128) {
        const int n = i / p.K, k = i % p.K;
        const float v = p.B[i];
        const float hi = p.nsplit == 3 ? __uint_as_float(__float_as_uint(v) & 0xFFFFE000u) : v;
        const int o = p.b_mn ? off_mnmajor(n, k, p.N) : off_kmajor(n, k, p.K);
        *(float *)(Bhi + o) = hi;
        *(float *)(Blo + o) = v - hi;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> async proxy (UMMA)
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                               ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(p.M >> 4) << 24);
        // per-operand descriptor fields and the start-address step of one K = 8 instruction
        const uint32_t a_lbo = p.a_mn ? (p.M / 4) * 128 : 128, a_sbo = p.a_mn ? 128 : (p.K / 4) * 128;
        const uint32_t b_lbo = p.b_mn ? (p.N / 4) * 128 : 128, b_sbo = p.b_mn ? 128 : (p.K / 4) * 128;
        const uint32_t a_step = p.a_mn ? (p.M / 4) * 128 : 256, b_step = p.b_mn ? (p.N / 4) * 128 : 256;
        int first = 1;
        for (int s = 0; s < p.nsplit; ++s) {
            // 3xTF32: hi*hi + lo*hi + hi*lo (small terms first would be better; order kept simple)
            const unsigned char *Ab = (s == 1) ? Alo : Ahi, *Bb = (s == 2) ? Blo : Bhi;
            for (int ks = 0; ks < p.K / 8; ++ks) {
                const uint64_t ad = make_desc(smem_u32(Ab) + ks * a_step, a_lbo, a_sbo);
                const uint64_t bd = make_desc(smem_u32(Bb) + ks * b_step, b_lbo, b_sbo);
                const uint32_t acc = first ? 0u : 1u;
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
                first = 0;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {   // everyone waits for the MMAs
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < p.N; c0 += 8) {
        uint32_t v[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) p.D[(warp * 32 + lane) * p.N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

static float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
static float tf32_rn(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x00000FFFu + ((u >> 13) & 1u); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

int main() {
    CHECK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const int K = 32;
    float *dA, *dB, *dD;
    CHECK(cudaMalloc(&dA, 128 * K * 4)); CHECK(cudaMalloc(&dB, 64 * K * 4)); CHECK(cudaMalloc(&dD, 128 * 64 * 4));
    static float A[128 * K], B[64 * K], D[128 * 64];
    for (int Mi = 0; Mi < 2; ++Mi) for (int Ni = 0; Ni < 2; ++Ni) for (int amn = 0; amn < 2; ++amn) for (int bmn = 0; bmn < 2; ++bmn) {
        const int M = Mi ? 128 : 64, N = Ni ? 64 : 32;
        // small integers: exact in tf32, every (m, n) result distinct enough to locate
        for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) A[m * K + k] = (float)((m * 37 + k * 11 + (m * k) % 7 + (m / 9) * 5) % 15 - 7);
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[n * K + k] = (float)((n * 29 + k * 13 + (n * k) % 5) % 15 - 7);
        CHECK(cudaMemcpy(dA, A, M * K * 4, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dB, B, N * K * 4, cudaMemcpyHostToDevice));
        CHECK(cudaMemset(dD, 0xFF, 128 * 64 * 4));
        Params p = {M, N, K, amn, bmn, 1, dA, dB, dD};
        probe<<<1, 128, 2 * (M + N) * K * 4>>>(p);
        CHECK(cudaDeviceSynchronize());
        CHECK(cudaMemcpy(D, dD, 128 * N * 4, cudaMemcpyDeviceToHost));
        // locate each accumulator row
        int bad = 0, lane_of_row[128];
        for (int m = 0; m < M; ++m) {
            lane_of_row[m] = -1;
            for (int l = 0; l < 128 && lane_of_row[m] < 0; ++l) {
                int ok = 1;
                for (int n = 0; n < N && ok; ++n) {
                    double ref = 0; for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * B[n * K + k];
                    ok = D[l * N + n] == (float)ref;
                }
                if (ok) lane_of_row[m] = l;
            }
            bad += lane_of_row[m] < 0;
        }
        printf("M=%3d N=%2d A-%s B-%s : %s rows found %d/%d; row->lane:", M, N, amn ? "MN" : "K ", bmn ? "MN" : "K ", bad ? "MISMATCH" : "ok", M - bad, M);
        for (int m = 0; m < M; m += 8) printf(" %d:%d", m, lane_of_row[m]);
        printf("\n");
    }
    {   // truncation vs rounding of fp32 inputs, and the 3xTF32 split
        const int M = 64, N = 32;
        srand(1);
        for (int i = 0; i < M * K; ++i) A[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
        for (int i = 0; i < N * K; ++i) B[i] = (float)rand() / RAND_MAX * 2.f - 1.f;
        CHECK(cudaMemcpy(dA, A, M * K * 4, cudaMemcpyHostToDevice)); CHECK(cudaMemcpy(dB, B, N * K * 4, cudaMemcpyHostToDevice));
        for (int nsplit = 1; nsplit <= 3; nsplit += 2) {
            Params p = {M, N, K, 0, 0, nsplit, dA, dB, dD};
            probe<<<1, 128, 2 * (M + N) * K * 4>>>(p);
            CHECK(cudaDeviceSynchronize());
            CHECK(cudaMemcpy(D, dD, 128 * N * 4, cudaMemcpyDeviceToHost));
            double e_exact = 0, e_trunc = 0, e_rn = 0, scale = 0;
            for (int m = 0; m < 16; ++m) for (int n = 0; n < N; ++n) {          // rows 0..15 sit in lanes 0..15 for M = 64 and 128
                double ex = 0, tr = 0, rn = 0, sa = 0;
                for (int k = 0; k < K; ++k) {
                    ex += (double)A[m * K + k] * B[n * K + k];
                    tr += (double)tf32_trunc(A[m * K + k]) * tf32_trunc(B[n * K + k]);
                    rn += (double)tf32_rn(A[m * K + k]) * tf32_rn(B[n * K + k]);
                    sa += fabs((double)A[m * K + k] * B[n * K + k]);
                }
                const double got = D[m * N + n];
                e_exact = fmax(e_exact, fabs(got - ex)); e_trunc = fmax(e_trunc, fabs(got - tr)); e_rn = fmax(e_rn, fabs(got - rn));
                scale = fmax(scale, sa);
            }
            printf("nsplit=%d: max|D-exact|=%.3e  max|D-trunc model|=%.3e  max|D-rn model|=%.3e  (sum|terms| up to %.2f)\n",
                   nsplit, e_exact, e_trunc, e_rn, scale);
        }
    }
    return 0;
}

// tcgen05 eval kernel of libb200env (BASELINE config 4 / 5: MLP 784 -> 64 -> 10, minibatch 32).
//
// One env-eval = loss + batch-SUM gradient of the classifier at the env's current parameters on its
// current minibatch (reference problems/optimize_nn.py:35-52, 122-159): forward Hpre = X.W1
// ([32 x D].[D x 64]), the small tail (bias, relu, second layer, softmax cross-entropy and their
// backward), backward G1 = X^T.dPre ([D x 32].[32 x 64]).  The two GEMMs run on the tensor cores as
// tcgen05.mma kind::tf32 with a 3xTF32 split that keeps the fp32 parity bar: the tensor core
// truncates an fp32 word to tf32, so the raw data IS the "hi" operand and only
// lo = rn_tf32(x - trunc(x)) is computed by threads; hi and lo of both operands are stacked along M
// and N, so one MMA per K step yields all four split products (lo.lo included: with truncation the
// lo parts are one-sided and their product is not negligible).
//
// Persistent, one 512-thread CTA per SM, warp-specialised; envs e = blockIdx.x + k * gridDim.x:
//   warp 0      W producer  : W1 tile [32 f][64 j] of the forward unit -> stage, 16-byte cp.async
//   warp 1      XF producer : minibatch rows [32 s][32 f] of the forward unit (gather), cp.async
//   warp 2      XB producer : minibatch rows [32 s][128 f] of a backward tile (L2 hits), cp.async
//   warp 3      MMA issuer  : one elected lane; forward of env k+1 is issued BEFORE backward of env k,
//                             so the tail of env k runs under the forward loads / MMAs of env k+1
//   warps 4-7   tail        : TMEM -> Hpre, bias/relu/layer 2/softmax-CE/backward -> dPre operand
//   warps 8-11  drain       : gradient tiles TMEM -> registers -> shared-memory transpose -> HBM;
//                             second eval: the step's scalars (reward, done, info, cursor)
//   warps 12-15 converters  : lo parts of every landed operand tile
// Pipelines: forward ring (4 stages x 24 KB: W hi/lo, X hi/lo), backward ring (2 x 32 KB: X hi/lo),
// both full -> converted -> (tcgen05.commit) empty; TMEM: four forward accumulators [128 x 64] (every
// fourth unit each: the tensor core's adder truncates, so long sums are split and added by threads) and
// two gradient accumulators [128 x 128], with full / free mbarrier pairs.
//
// Shared-memory operand layouts (UMMA canonical, 128-byte rows):
//   MN-major tf32 (W1 tile as A, X^T as A, dPre as B): blocks [K rows][32 MN elements], 32-byte
//     chunks XORed with (row & 3) (SWIZZLE_128B_BASE32B, the only MN-major layout tf32 has),
//     SBO = 512 (4-row groups), LBO = 4096 (next 32 MN elements)
//   K-major tf32 (X as B of the forward): rows [N][32 K elements], 16-byte chunks XORed with
//     (row & 7) (SWIZZLE_128B), SBO = 1024
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "b200env_shared.cuh"
#include "b200tc.h"

namespace {
namespace tc2 {

constexpr int N1 = 64, B = 32;
constexpr int SF = 4, SB = 2;
constexpr int F_STAGE = 24576, F_W = 0, F_WLO = 8192, F_X = 16384, F_XLO = 20480;
constexpr int B_STAGE = 32768, B_XLO = 16384;
constexpr int OFF_F = 0;
constexpr int OFF_B = OFF_F + SF * F_STAGE;
constexpr int OFF_DP = OFF_B + SB * B_STAGE;          // dPre operand: hi j0, hi j1, lo j0, lo j1 (4 KB each)
constexpr int OFF_G = OFF_DP + 16384;                 // gradient staging: 4 warps x 4 KB
constexpr int OFF_TAIL = OFF_G + 16384;               // float scratch of the tail
// tail scratch (floats)
constexpr int CMAX = 16;
constexpr int TW_MAX = N1 + N1 * CMAX + CMAX;         // b1, W2, b2
constexpr int T_H = 0, T_TW = T_H + B * N1, T_TG = T_TW + TW_MAX, T_Z = T_TG + TW_MAX, T_LB = T_Z + B * CMAX,
              T_YS = T_LB + B, T_GP = T_YS + B, T_MISC = T_GP + 2 * N1, T_END = T_MISC + 16;
constexpr int SMEM_BYTES = OFF_TAIL + T_END * 4 + 1024;   // + slack to align the base to 1024 bytes
constexpr int TMEM_COLS = 512;                        // forward 4 x 64, gradient 2 x 128 columns
constexpr int TM_F = 0, TM_G = 256, NACC = 4;
constexpr int THREADS = 512;
constexpr long long WATCHDOG_CYCLES = 1500000000LL;   // ~0.8 s: a wait this long is a protocol bug

struct Bars {
    uint64_t fullF[SF], convF[SF], emptyF[SF];
    uint64_t fullB[SB], convB[SB], emptyB[SB];
    uint64_t g_full[2], g_free[2], tail_done[2];
    uint64_t fwd_done, fwd_free, dpre_ready, dpre_free;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a protocol error ends the kernel with a code in dbg[] instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int code, int *dbg) {
    uint32_t done = 0;
    int spins = 0;
    long long t0 = 0;
    while (true) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
        if (++spins == 64) t0 = clock64();
        if (spins > 64 && (spins & 63) == 0) {
            if (*reinterpret_cast<volatile int *>(dbg) != 0) return;          // somebody else gave up
            if (clock64() - t0 > WATCHDOG_CYCLES) {
                if (atomicCAS(dbg, 0, code) == 0) { dbg[1] = (int)blockIdx.x; dbg[2] = (int)threadIdx.x; dbg[3] = (int)parity; }
                return;
            }
        }
    }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {       // arrive when this thread's copies have landed
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t type) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46) | ((uint64_t)type << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void group_bar(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }

#define TC2_LD32(taddr, v) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
                 : "r"(taddr))

// lo part of the 3xTF32 split: the tensor core sees trunc(x) of the raw word, lo carries the rest,
// itself rounded to nearest at tf32 precision (its own truncation would be one-sided)
__device__ __forceinline__ float lo_of(float x) {
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    const float lo = x - hi;
    return __uint_as_float((__float_as_uint(lo) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float4 lo_of4(float4 v) { return make_float4(lo_of(v.x), lo_of(v.y), lo_of(v.z), lo_of(v.w)); }
// byte offset of 16-byte chunk q (0..7) of 128-byte row r: 32-byte chunks ^ (r & 3)  /  16-byte chunks ^ (r & 7)
__device__ __forceinline__ uint32_t swz32(int q, int r) { return (uint32_t)(((((q >> 1) ^ (r & 3)) << 5) | ((q & 1) << 4))); }
__device__ __forceinline__ uint32_t swz16(int q, int r) { return (uint32_t)((q ^ (r & 7)) << 4); }

template <bool SECOND, int CC>
__global__ void __launch_bounds__(THREADS, 1) tc2_eval_kernel(const __grid_constant__ Dev d,
                                                              const __grid_constant__ StepArgs a, int *dbg) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) Bars bars;
    __shared__ uint32_t tmem_slot;
    __shared__ float loss_slot[2];
    __shared__ double gsum_slot[2];
    __shared__ double red_d[8];
    __shared__ float misc_s[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *const sm = smem_raw + (sbase - smem_u32(smem_raw));
    float *const ts = reinterpret_cast<float *>(sm + OFF_TAIL);
    const int D = d.D, C = CC ? CC : d.C;
    const int UF = (D + 31) >> 5, TB = (D + 127) >> 7;
    const int e_end = a.e_begin + a.e_count;

    if (tid == 0) {
        for (int s = 0; s < SF; ++s) { mbar_init(&bars.fullF[s], 64); mbar_init(&bars.convF[s], 4); mbar_init(&bars.emptyF[s], 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(&bars.fullB[s], 32); mbar_init(&bars.convB[s], 4); mbar_init(&bars.emptyB[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars.g_full[s], 1); mbar_init(&bars.g_free[s], 4); mbar_init(&bars.tail_done[s], 1);
        }
        mbar_init(&bars.fwd_done, 1); mbar_init(&bars.fwd_free, 4);
        mbar_init(&bars.dpre_ready, 1); mbar_init(&bars.dpre_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        // ===================== W producer: forward unit u = rows f in [32u, 32u + 32) of W1, both column halves
        uint32_t it = 0;
        for (int e = a.e_begin + blockIdx.x; e < e_end; e += gridDim.x) {
            const float *We = d.w + (size_t)e * d.Pp;
            for (int u = 0; u < UF; ++u, ++it) {
                const int s = it % SF;
                mbar_wait(&bars.emptyF[s], ((it / SF) & 1) ^ 1, 1, dbg);
                const uint32_t base = sbase + OFF_F + s * F_STAGE + F_W;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int c = i * 32 + lane, r = c >> 4, q = c & 15, f = 32 * u + r;
                    cp_async16(base + (q >> 3) * 4096 + r * 128 + swz32(q & 7, r),
                               We + (size_t)(f < D ? f : 0) * N1 + 4 * q, f < D ? 16 : 0);
                }
                cp_async_arrive(&bars.fullF[s]);
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp == 1) {
        // ===================== XF producer: forward unit u = features [32u, 32u + 32) of the 32 minibatch rows
        uint32_t it = 0;
        for (int e = a.e_begin + blockIdx.x; e < e_end; e += gridDim.x) {
            const int *idx; int cnt;
            current_batch(d, a, e, d.sc + e, idx, cnt);
            const int my_row = lane < cnt ? idx[lane] : -1;
            for (int u = 0; u < UF; ++u, ++it) {
                const int s = it % SF;
                mbar_wait(&bars.emptyF[s], ((it / SF) & 1) ^ 1, 2, dbg);
                const uint32_t base = sbase + OFF_F + s * F_STAGE + F_X;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int sr = i * 4 + (lane >> 3), q = lane & 7, f = 32 * u + 4 * q;
                    const int row = __shfl_sync(0xffffffffu, my_row, sr);
                    const bool ok = row >= 0 && f < D;
                    cp_async16(base + sr * 128 + swz16(q, sr), d.X + (ok ? (size_t)row * d.Dp + f : 0), ok ? 16 : 0);
                }
                cp_async_arrive(&bars.fullF[s]);
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp == 2) {
        // ===================== XB producer: backward tile t = features [128t, 128t + 128), four 32-feature blocks
        uint32_t it = 0;
        for (int e = a.e_begin + blockIdx.x; e < e_end; e += gridDim.x) {
            const int *idx; int cnt;
            current_batch(d, a, e, d.sc + e, idx, cnt);
            const int my_row = lane < cnt ? idx[lane] : -1;
            for (int t = 0; t < TB; ++t, ++it) {
                const int s = it % SB;
                mbar_wait(&bars.emptyB[s], ((it / SB) & 1) ^ 1, 3, dbg);
                const uint32_t base = sbase + OFF_B + s * B_STAGE;
#pragma unroll 4
                for (int i = 0; i < 32; ++i) {
                    const int blk = i >> 3, sr = (i & 7) * 4 + (lane >> 3), q = lane & 7, f = 128 * t + 32 * blk + 4 * q;
                    const int row = __shfl_sync(0xffffffffu, my_row, sr);
                    const bool ok = row >= 0 && f < D;
                    cp_async16(base + blk * 4096 + sr * 128 + swz32(q, sr), d.X + (ok ? (size_t)row * d.Dp + f : 0), ok ? 16 : 0);
                }
                cp_async_arrive(&bars.fullB[s]);
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp == 3) {
        // ===================== MMA issuer
        constexpr uint32_t TF32 = (1u << 4) | (2u << 7) | (2u << 10);
        const uint32_t idesc_f = TF32 | (1u << 15) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);               // A MN-major, B K-major
        const uint32_t idesc_b = TF32 | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); // both MN-major
        uint32_t itF = 0, itB = 0, itG = 0;
        const uint64_t dpd = make_desc(sbase + OFF_DP, 4096, 512, 1);
        auto backward = [&](int j) {                     // j = local index of the env whose dPre is ready
            mbar_wait(&bars.dpre_ready, j & 1, 4, dbg);
            for (int t = 0; t < TB; ++t, ++itB, ++itG) {
                const int s = itB % SB, g = itG & 1;
                mbar_wait(&bars.convB[s], (itB / SB) & 1, 5, dbg);
                mbar_wait(&bars.g_free[g], ((itG >> 1) & 1) ^ 1, 6, dbg);
                fence_async();
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t ahi = make_desc(sbase + OFF_B + s * B_STAGE, 4096, 512, 1);
                    const uint64_t alo = make_desc(sbase + OFF_B + s * B_STAGE + B_XLO, 4096, 512, 1);
                    const uint32_t acc = tmem + TM_G + 128 * g;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {     // 8 samples per MMA = 1024 bytes of each block
                        mma_tf32(acc, ahi + (uint64_t)(kk * 64), dpd + (uint64_t)(kk * 64), idesc_b, kk ? 1u : 0u);
                        mma_tf32(acc, alo + (uint64_t)(kk * 64), dpd + (uint64_t)(kk * 64), idesc_b, 1u);
                    }
                    mma_commit(&bars.emptyB[s]);
                    mma_commit(&bars.g_full[g]);
                    if (t == TB - 1) mma_commit(&bars.dpre_free);
                }
                __syncwarp();
            }
        };
        int k = 0;
        for (int e = a.e_begin + blockIdx.x; e < e_end; e += gridDim.x, ++k) {
            mbar_wait(&bars.fwd_free, (k & 1) ^ 1, 7, dbg);       // the tail has read the accumulators of env k-1
            for (int u = 0; u < UF; ++u, ++itF) {
                const int s = itF % SF;
                mbar_wait(&bars.convF[s], (itF / SF) & 1, 8, dbg);
                fence_async();
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t ad = make_desc(sbase + OFF_F + s * F_STAGE + F_W, 4096, 512, 1);
                    const uint64_t bd = make_desc(sbase + OFF_F + s * F_STAGE + F_X, 16, 1024, 2);
                    // The tensor core's adder truncates: a sum over all 98 K steps in ONE accumulator
                    // drifts by ~1e-6 of the result.  Four accumulators take every fourth unit
                    // (<= 28 steps each) and the tail adds them in fp32 round-to-nearest.
                    const uint32_t acc = tmem + TM_F + 64 * (u & (NACC - 1));
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)       // 8 features per MMA: 1024 bytes of A, 32 bytes of B
                        mma_tf32(acc, ad + (uint64_t)(kk * 64), bd + (uint64_t)(kk * 2), idesc_f, (u >= NACC || kk) ? 1u : 0u);
                    mma_commit(&bars.emptyF[s]);
                    if (u == UF - 1) mma_commit(&bars.fwd_done);
                }
                __syncwarp();
            }
            if (k >= 1) backward(k - 1);
        }
        if (k >= 1) backward(k - 1);
    } else if (warp < 8) {
        // ===================== tail group (128 threads, named barrier 1)
        const int q = warp - 4, ttid = tid - 128;
        float *Hb = ts + T_H, *tw = ts + T_TW, *tg = ts + T_TG, *Z = ts + T_Z, *lb = ts + T_LB, *gp = ts + T_GP;
        int *ys = reinterpret_cast<int *>(ts + T_YS);
        float *tmisc = ts + T_MISC;
        const int Zs = CMAX;
        const int tailP = d.tailP;
        const float *b1 = tw, *W2 = tw + N1, *b2 = tw + N1 + N1 * C;
        unsigned char *dP = sm + OFF_DP;
        int k = 0;
        for (int e = a.e_begin + blockIdx.x; e < e_end; e += gridDim.x, ++k) {
            const float *We = d.w + (size_t)e * d.Pp;
            float *gout = d.gnext + (size_t)e * d.Pp;
            const int *idx; int cnt;
            current_batch(d, a, e, d.sc + e, idx, cnt);
            for (int i = ttid; i < tailP; i += 128) tw[i] = We[d.P1 + i];
            if (ttid < B) ys[ttid] = ttid < cnt ? d.labels[idx[ttid]] : 0;
            // ---- forward accumulator: lanes 0..63 = W_hi rows, 64..127 = W_lo rows; columns 0..31 = X_hi, 32..63 = X_lo
            mbar_wait(&bars.fwd_done, k & 1, 9, dbg);
            tc_fence_after();
            float acc[32];
            {
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + TM_F;
                const int nacc = UF < NACC ? UF : NACC;
                float lo_part[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) { acc[i] = 0.f; lo_part[i] = 0.f; }
                for (int c = 0; c < nacc; ++c) {
                    uint32_t v0[32], v1[32];
                    TC2_LD32(taddr + 64 * c + 32, v1);
                    TC2_LD32(taddr + 64 * c, v0);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int i = 0; i < 32; ++i) { lo_part[i] += __uint_as_float(v1[i]); acc[i] += __uint_as_float(v0[i]); }
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] += lo_part[i];
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.fwd_free);
            const int j = (q & 1) * 32 + lane;
            if (q >= 2) {
#pragma unroll
                for (int s = 0; s < B; ++s) Hb[s * N1 + j] = acc[s];
            }
            group_bar(1);
            if (q < 2) {                                  // + bias, relu (rows the minibatch does not have stay 0)
                const float bj = b1[j];
#pragma unroll
                for (int s = 0; s < B; ++s) {
                    const float v = (Hb[s * N1 + j] + acc[s]) + bj;
                    Hb[s * N1 + j] = (s < cnt && v > 0.f) ? v : 0.f;
                }
            }
            group_bar(1);
            // ---- second layer, softmax cross-entropy (optimize_nn.py:42-50)
            for (int i = ttid; i < cnt * C; i += 128) {
                const int s = i / C, c = i - s * C;
                float z = b2[c];
#pragma unroll 8
                for (int jj = 0; jj < N1; ++jj) z = fmaf(Hb[s * N1 + jj], W2[jj * C + c], z);
                Z[s * Zs + c] = z;
            }
            group_bar(1);
            if (ttid < B) {
                const int s = ttid;
                float loss = 0.f;
                float *z = Z + s * Zs;
                if (s < cnt) {
                    const int y = ys[s];
                    float m = z[0];
                    for (int c = 1; c < C; ++c) m = fmaxf(m, z[c]);
                    float sum = 0.f;
                    for (int c = 0; c < C; ++c) sum += expf(z[c] - m);
                    const float zy = z[y];
                    loss = (m + logf(sum)) - zy;
                    const float inv = 1.0f / sum;
                    for (int c = 0; c < C; ++c) {
                        const float p = expf(z[c] - m) * inv;
                        z[c] = p - (c == y ? 1.f : 0.f);
                    }
                } else {
                    for (int c = 0; c < C; ++c) z[c] = 0.f;
                }
                lb[s] = loss;
            }
            group_bar(1);
            // ---- backward of the tail: gW2, gb2, dPre (and gb1 from its column sums)
            float *gb1 = tg, *gW2 = tg + N1, *gb2 = tg + N1 + N1 * C;
            if (ttid == 0) {
                float l = 0.f;
                for (int s = 0; s < cnt; ++s) l += lb[s];
                tmisc[0] = l / (float)cnt;
            }
            for (int i = ttid; i < N1 * C; i += 128) {
                const int jj = i / C, c = i - jj * C;
                float g = 0.f;
                for (int s = 0; s < cnt; ++s) g = fmaf(Hb[s * N1 + jj], Z[s * Zs + c], g);
                gW2[i] = g;
            }
            if (ttid < C) {
                float g = 0.f;
                for (int s = 0; s < cnt; ++s) g += Z[s * Zs + ttid];
                gb2[ttid] = g;
            }
            if (k >= 1) mbar_wait(&bars.dpre_free, (k - 1) & 1, 10, dbg);   // backward MMAs of the previous env have read dPre
            {
                const int jj = ttid & 63, sg = ttid >> 6;
                float w2r[CMAX];
#pragma unroll
                for (int c = 0; c < CMAX; ++c) w2r[c] = c < C ? W2[jj * C + c] : 0.f;
                float colsum = 0.f;
                const uint32_t cbase = (uint32_t)((jj >> 5) * 4096 + ((jj & 7) << 2));
                const int q32 = (jj & 31) >> 3;
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const int s = sg * 16 + i;
                    float v = 0.f;
                    if (s < cnt && Hb[s * N1 + jj] > 0.f) {
#pragma unroll
                        for (int c = 0; c < CMAX; ++c)
                            if (c < C) v = fmaf(Z[s * Zs + c], w2r[c], v);
                    }
                    colsum += v;
                    const uint32_t off = cbase + (uint32_t)(s * 128) + (uint32_t)((q32 ^ (s & 3)) << 5);
                    *reinterpret_cast<float *>(dP + off) = v;
                    *reinterpret_cast<float *>(dP + 8192 + off) = lo_of(v);
                }
                gp[sg * N1 + jj] = colsum;
            }
            fence_async();
            group_bar(1);
            if (ttid == 0) mbar_arrive(&bars.dpre_ready);
            if (ttid < N1) gb1[ttid] = gp[ttid] + gp[N1 + ttid];
            group_bar(1);
            // ---- tail gradient to HBM
            float gsum = 0.f;
            for (int i = ttid; i < tailP; i += 128) {
                const float g = tg[i];
                gout[d.P1 + i] = g;
                gsum += g;
            }
            if (SECOND) {
                double v = warp_sum((double)gsum);
                if (lane == 0) red_d[q] = v;
                group_bar(1);
                if (ttid == 0) {
                    gsum_slot[k & 1] = (red_d[0] + red_d[1]) + (red_d[2] + red_d[3]);
                    loss_slot[k & 1] = tmisc[0];
                    __threadfence_block();
                    mbar_arrive(&bars.tail_done[k & 1]);
                }
            } else if (a.loss_out != nullptr && ttid == 0) {
                a.loss_out[e] = tmisc[0];
            }
            group_bar(1);                                 // scratch is reused by the next env
        }
    } else if (warp < 12) {
        // ===================== drain group (128 threads, named barrier 2): gradient tiles -> HBM
        const int q = warp - 8, dtid = tid - 256;
        unsigned char *stg = sm + OFF_G + q * 4096;
        uint32_t itG = 0;
        int k = 0;
        for (int e = a.e_begin + blockIdx.x; e < e_end; e += gridDim.x, ++k) {
            float *gout = d.gnext + (size_t)e * d.Pp;
            float gsum = 0.f;
            for (int t = 0; t < TB; ++t, ++itG) {
                const int g = itG & 1;
                mbar_wait(&bars.g_full[g], (itG >> 1) & 1, 11, dbg);
                tc_fence_after();
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + TM_G + 128 * g;
                const int f0 = 128 * t + 32 * q;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float v[32];
                    {
                        uint32_t v0[32], v1[32];
                        TC2_LD32(taddr + 64 + 32 * h, v1);            // x . dPre_lo
                        TC2_LD32(taddr + 32 * h, v0);                 // x . dPre_hi
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(v1[i]) + __uint_as_float(v0[i]);
                    }
                    if (h == 1) {                                     // accumulator drained: the next tile may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars.g_free[g]);
                    }
                    // lane = feature row; transpose through the warp's staging block so that stores are full lines
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        *reinterpret_cast<float4 *>(stg + lane * 128 + swz16(c, lane)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                        gsum += (v[4 * c] + v[4 * c + 1]) + (v[4 * c + 2] + v[4 * c + 3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = i * 4 + (lane >> 3), c = lane & 7;
                        const float4 o = *reinterpret_cast<const float4 *>(stg + r * 128 + swz16(c, r));
                        if (f0 + r < D) *reinterpret_cast<float4 *>(gout + (size_t)(f0 + r) * N1 + 32 * h + 4 * c) = o;
                    }
                }
            }
            if (!SECOND) continue;
            // ---- scalars of the step: history bookkeeping, reward, done, info (multioptlrs.py:89-128)
            double v = warp_sum((double)gsum);
            if (lane == 0) red_d[4 + q] = v;
            group_bar(2);
            if (dtid == 0) {
                mbar_wait(&bars.tail_done[k & 1], (k >> 1) & 1, 12, dbg);
                const double gtot = ((red_d[4] + red_d[5]) + (red_d[6] + red_d[7])) + gsum_slot[k & 1];
                step_scalars(d, a, d.sc + e, e, loss_slot[k & 1], gtot, misc_s);
            }
            group_bar(2);
            if (misc_s[4] != 0.f) {                       // epoch wrapped: InMemoryDataSet.on_epoch_end with the env's permutation
                EnvScalars *sc = d.sc + e;
                const int sel = sc->ord_sel;
                const int *src = order_ptr(d, e, sel);
                int *dst = d.ord + ((size_t)(sel ^ 1) * d.E + e) * d.N;
                const int *pm = d.perm + (size_t)e * d.perm_stride;
                for (int i = dtid; i < d.N; i += 128) dst[i] = src[pm[i]];
                group_bar(2);
                if (dtid == 0) sc->ord_sel = sel ^ 1;
            }
            group_bar(2);
        }
    } else {
        // ===================== converters (128 threads): lo parts of the landed tiles
        const int ctid = tid - 384;
        uint32_t itF = 0, itB = 0;
        auto convert_b = [&]() {
            for (int t = 0; t < TB; ++t, ++itB) {
                const int s = itB % SB;
                mbar_wait(&bars.fullB[s], (itB / SB) & 1, 13, dbg);
                unsigned char *base = sm + OFF_B + s * B_STAGE;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int off = (i * 128 + ctid) * 16;
                    *reinterpret_cast<float4 *>(base + B_XLO + off) = lo_of4(*reinterpret_cast<const float4 *>(base + off));
                }
                fence_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.convB[s]);
            }
        };
        int k = 0;
        for (int e = a.e_begin + blockIdx.x; e < e_end; e += gridDim.x, ++k) {
            for (int u = 0; u < UF; ++u, ++itF) {
                const int s = itF % SF;
                mbar_wait(&bars.fullF[s], (itF / SF) & 1, 14, dbg);
                unsigned char *base = sm + OFF_F + s * F_STAGE;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int off = (i * 128 + ctid) * 16;
                    *reinterpret_cast<float4 *>(base + F_WLO + off) = lo_of4(*reinterpret_cast<const float4 *>(base + F_W + off));
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int off = (i * 128 + ctid) * 16;
                    *reinterpret_cast<float4 *>(base + F_XLO + off) = lo_of4(*reinterpret_cast<const float4 *>(base + F_X + off));
                }
                fence_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.convF[s]);
            }
            if (k >= 1) convert_b();
        }
        if (k >= 1) convert_b();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace tc2
}  // namespace

// ---------------------------------------------------------------- host side
bool b2e_tc2_supported(const void *dev) {
    const Dev &d = *static_cast<const Dev *>(dev);
    return d.kind == B2E_PROBLEM_SOFTMAX && d.hidden && !d.generic && d.N1 == tc2::N1 && d.B == tc2::B &&
           d.C >= 1 && d.C <= tc2::CMAX && d.D >= 32 && d.D % 4 == 0 && d.Dp == d.D && d.Pp % 4 == 0;
}

size_t b2e_tc2_smem_bytes() { return (size_t)tc2::SMEM_BYTES; }

const char *b2e_tc2_prepare() {
    const void *fns[] = {(const void *)tc2::tc2_eval_kernel<false, 10>, (const void *)tc2::tc2_eval_kernel<true, 10>,
                         (const void *)tc2::tc2_eval_kernel<false, 0>, (const void *)tc2::tc2_eval_kernel<true, 0>};
    for (const void *fn : fns)
        if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES) != cudaSuccess)
            return "tcgen05 eval kernel does not fit shared memory";
    return nullptr;
}

int b2e_tc2_launch(const void *dev, const void *args, int second, int grid, int *dbg, void *stream) {
    const Dev &d = *static_cast<const Dev *>(dev);
    const StepArgs &a = *static_cast<const StepArgs *>(args);
    const cudaStream_t cs = (cudaStream_t)stream;
    if (d.C == 10) {
        if (second) tc2::tc2_eval_kernel<true, 10><<<grid, tc2::THREADS, tc2::SMEM_BYTES, cs>>>(d, a, dbg);
        else tc2::tc2_eval_kernel<false, 10><<<grid, tc2::THREADS, tc2::SMEM_BYTES, cs>>>(d, a, dbg);
    } else {
        if (second) tc2::tc2_eval_kernel<true, 0><<<grid, tc2::THREADS, tc2::SMEM_BYTES, cs>>>(d, a, dbg);
        else tc2::tc2_eval_kernel<false, 0><<<grid, tc2::THREADS, tc2::SMEM_BYTES, cs>>>(d, a, dbg);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

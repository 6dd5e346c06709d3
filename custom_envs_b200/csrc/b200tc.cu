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
// Persistent, one 768-thread CTA per SM, warp-specialised; envs e = blockIdx.x + k * gridDim.x:
//   warp 0      W producer  : W1 tile [32 f][64 j] of the forward unit -> stage, one bulk tensor copy; L2
//                             prefetch of the next env's parameters
//   warp 1      XF producer : minibatch rows [32 s][32 f] of the forward unit (gather), 16-byte cp.async
//   warp 2      XB producer : minibatch rows [32 s][128 f] of a backward tile (L2 hits), 16-byte cp.async
//   warps 3, 23 MMA issuers : forward units / backward tiles, one elected lane each, two independent
//                             in-order queues; the tail of env k runs under the forward stream of env k+1
//   warps 4-7   tail        : Hpre -> bias/relu/layer 2/softmax-CE/backward -> dPre operand, tail gradient
//   warps 8-11  drain       : gradient tiles TMEM -> registers -> shared-memory transpose -> HBM;
//                             second eval: the step's scalars (reward, done, info, cursor)
//   warps 12-15, 20-22 converters : lo parts of every landed operand tile (forward ring / backward ring)
//   warps 16-19 read-out    : forward accumulators TMEM -> fp32 registers every two units (short sums inside
//                             the truncating tensor-core adder), Hpre of the env -> tail group
// Pipelines: forward ring (4 stages x 24 KB: W hi/lo, X hi/lo), backward ring (2 x 32 KB: X hi/lo),
// both full -> converted -> (tcgen05.commit) empty; TMEM: four forward accumulators [128 x 64] (two
// units each, then read out: the tensor core's adder truncates, so long sums are added by threads) and
// two gradient accumulators [128 x 128], with full / free mbarrier pairs.
//
// Shared-memory operand layouts (UMMA canonical, 128-byte rows):
//   MN-major tf32 (W1 tile as A, X^T as A, dPre as B): blocks [K rows][32 MN elements], 32-byte
//     chunks XORed with (row & 3) (SWIZZLE_128B_BASE32B, the only MN-major layout tf32 has),
//     SBO = 512 (4-row groups), LBO = 4096 (next 32 MN elements)
//   K-major tf32 (X as B of the forward): rows [N][32 K elements], 16-byte chunks XORed with
//     (row & 7) (SWIZZLE_128B), SBO = 1024
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

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
constexpr int HS = B + 1;                             // row stride of the transposed activations H^T [64 j][32 s]
constexpr int T_H = 0, T_TW = T_H + N1 * HS, T_TG = T_TW + TW_MAX, T_ZP = T_TG + TW_MAX, T_Z = T_ZP + 4 * CMAX * B,
              T_YS = T_Z + B * CMAX, T_GP = T_YS + B, T_MISC = T_GP + 2 * N1, T_END = T_MISC + 16;
constexpr int SMEM_BYTES = OFF_TAIL + T_END * 4 + 1024;   // + slack to align the base to 1024 bytes
constexpr int TMEM_COLS = 512;                        // forward 4 x 64, gradient 2 x 128 columns
constexpr int TM_F = 0, TM_G = 256, NACC = 4, GRP = 1;     // forward accumulators; units summed inside the tensor core per read-out
constexpr int THREADS = 768;                          // 24 warps, at most 80 registers each
constexpr int CONV_F_WARPS = 4, READ_WARP0 = 16, CONV_B_WARP0 = 20, CONV_B_WARPS = 3, MMA_B_WARP = 23;
constexpr int DBG_BYTES = 64 + 8 * 512 * 8;
constexpr long long WATCHDOG_CYCLES = 1500000000LL;   // ~0.8 s: a wait this long is a protocol bug

struct Bars {
    uint64_t fullF[SF], convF[SF], emptyF[SF];
    uint64_t fullB[SB], convB[SB], emptyB[SB];
    uint64_t g_full[2], g_free[2], tail_done[2];
    uint64_t acc_full[NACC], acc_free[NACC], hpre_ready, hpre_free, dpre_ready, dpre_free;
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
__device__ __forceinline__ void prefetch_l2(const void *src, int bytes) {              // bytes: multiple of 16
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int bytes) {      // bytes < 16: zero fill
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

#define TC2_LD16(taddr, v) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
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
                                                              const __grid_constant__ StepArgs a,
                                                              const __grid_constant__ CUtensorMap map_w,
                                                              const __grid_constant__ CUtensorMap map_g0,
                                                              const __grid_constant__ CUtensorMap map_g1, int gsel, int pf_units, int *dbg) {
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
    // work items: the envs of the range, or of the list a reset built (b200env.cu reset_list_kernel)
    const int n_items = a.env_list ? *a.env_count : a.e_count;
    auto env_at = [&](int i) { return a.env_list ? a.env_list[i] : a.e_begin + i; };
    // optional timeline of CTA 0 (B2E_TC_TRACE=<file>): clock64 of the n-th event of each role
    long long *const trace = (dbg[15] != 0 && blockIdx.x == 0) ? reinterpret_cast<long long *>(dbg + 16) : nullptr;
#define TC2_TRACE(role, n) do { if (trace && lane == 0 && (n) < 512) trace[(role) * 512 + (n)] = clock64(); } while (0)

    if (tid == 0) {
        for (int s = 0; s < SF; ++s) { mbar_init(&bars.fullF[s], 33); mbar_init(&bars.convF[s], CONV_F_WARPS); mbar_init(&bars.emptyF[s], 1); }
        for (int s = 0; s < SB; ++s) { mbar_init(&bars.fullB[s], 32); mbar_init(&bars.convB[s], CONV_B_WARPS); mbar_init(&bars.emptyB[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bars.g_full[s], 1); mbar_init(&bars.g_free[s], 4); mbar_init(&bars.tail_done[s], 1);
        }
        for (int s = 0; s < NACC; ++s) { mbar_init(&bars.acc_full[s], 1); mbar_init(&bars.acc_free[s], 4); }
        mbar_init(&bars.dpre_ready, 1); mbar_init(&bars.dpre_free, 1);
        mbar_init(&bars.hpre_ready, 2); mbar_init(&bars.hpre_free, 1);
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
        // ===================== W producer.  This SM's copy engine retires one bulk tensor copy per
        // ~400-600 cycles whatever its size up to 16 KB (profiles/tools/tma_rate_probe.cu), so the W1 tile
        // of a unit travels as ONE copy: a 4-d view [E][2 column halves][D][32] of the parameter array
        // (strides 256 B / 128 B / Pp*4) with box {32, 32, 2, 1} lands as the two MN-major blocks
        // [half][32 f][128 B] of the stage, swizzle 128B_ATOM_32B; rows beyond D arrive as zeros.
        // The next env's parameters are requested into L2 one env ahead.
        auto prefetch_env = [&](int e) {
            const char *base = reinterpret_cast<const char *>(d.w + (size_t)e * d.Pp);
            const int lines = (d.P * 4 + 127) >> 7;
            for (int i = lane; i < lines; i += 32)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)i * 128));
        };
        uint32_t it = 0;
        if ((int)blockIdx.x < n_items && pf_units > 0) prefetch_env(env_at(blockIdx.x));
        for (int it_e = blockIdx.x; it_e < n_items; it_e += gridDim.x) {
            const int e = env_at(it_e);
            if (it_e + (int)gridDim.x < n_items && pf_units > 0) prefetch_env(env_at(it_e + gridDim.x));
            for (int u = 0; u < UF; ++u, ++it) {
                const int s = it % SF;
                mbar_wait(&bars.emptyF[s], ((it / SF) & 1) ^ 1, 1, dbg);
                if (lane == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars.fullF[s])), "r"(8192u) : "memory");
                    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                                 ::"r"(sbase + OFF_F + s * F_STAGE + F_W), "l"(&map_w), "r"(0), "r"(32 * u), "r"(0), "r"(e),
                                   "r"(smem_u32(&bars.fullF[s])) : "memory");
                }
                TC2_TRACE(0, it);
            }
        }
    } else if (warp == 1 || warp == 2) {
        // ===================== X producers (gather of the minibatch rows, 16-byte cp.async that complete on
        // the stage's mbarrier: of everything tried -- one bulk tensor copy per row piece (~250 cycles each),
        // loads through the converters' registers -- this leaves the forward stream the least exposed):
        // warp 1, forward unit u = features [32u, 32u + 32): sample sr = 4 i + (lane >> 3), chunk q = lane & 7,
        //         K-major SWIZZLE_128B rows [32 s][128 B];
        // warp 2, backward tile t = features [128t, 128t + 128): four 32-feature blocks of [32 s][128 B],
        //         MN-major 128B_BASE32B (L2 hits: the forward pass read these rows microseconds ago).
        // Per env the eight row pointers of a lane are fixed; per unit only a byte offset is added.
        const bool fwd = warp == 1;
        const int g8 = lane >> 3, q = lane & 7;
        uint32_t lane_dst[2];                             // forward: the swizzle phase of row 4 i + g8 alternates with i
        lane_dst[0] = (uint32_t)(g8 * 128) + (fwd ? swz16(q, g8) : swz32(q, g8));
        lane_dst[1] = (uint32_t)(g8 * 128) + (fwd ? swz16(q, 4 + g8) : swz32(q, g8));
        uint32_t it = 0;
        // the row indices of an env sit behind two dependent global loads (cursor -> order -> row):
        // they are fetched one env ahead, so that no env starts with that latency
        auto fetch_row = [&](int item) {
            if (item >= n_items) return -1;
            const int e = env_at(item);
            const int *idx; int cnt;
            current_batch(d, a, e, d.sc + e, idx, cnt);
            return lane < cnt ? idx[lane] : -1;
        };
        int next_row = fetch_row(blockIdx.x);
        for (int it_e = blockIdx.x; it_e < n_items; it_e += gridDim.x) {
            const int my_row = next_row;
            next_row = fetch_row(it_e + gridDim.x);          // consumed at unit 8 / at the next env: the loads have time to land
            const char *rowp[8];
            int rbytes[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int row = __shfl_sync(0xffffffffu, my_row, 4 * i + g8);
                rowp[i] = reinterpret_cast<const char *>(d.X + (row >= 0 ? (size_t)row * d.Dp : 0)) + 16 * q;
                rbytes[i] = row >= 0 ? 16 : 0;
            }
            if (fwd) {
                for (int u = 0; u < UF; ++u, ++it) {
                    const int s = it % SF;
                    mbar_wait(&bars.emptyF[s], ((it / SF) & 1) ^ 1, 2, dbg);
                    const uint32_t base = sbase + OFF_F + s * F_STAGE + F_X;
                    const bool in = 32 * u + 4 * q < D;   // the last unit may be partial
                    if (u == 8 && next_row >= 0) prefetch_l2(d.X + (size_t)next_row * d.Dp, D * 4);   // the next env's rows into L2
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        cp_async16(base + lane_dst[i & 1] + 512 * i, in ? rowp[i] + 128 * u : rowp[i], in ? rbytes[i] : 0);
                    cp_async_arrive(&bars.fullF[s]);
                }
            } else {
                for (int t = 0; t < TB; ++t, ++it) {
                    const int s = it % SB;
                    mbar_wait(&bars.emptyB[s], ((it / SB) & 1) ^ 1, 3, dbg);
                    const uint32_t base = sbase + OFF_B + s * B_STAGE;
#pragma unroll
                    for (int blk = 0; blk < 4; ++blk) {
                        const bool in = 128 * t + 32 * blk + 4 * q < D;
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            cp_async16(base + lane_dst[0] + 4096 * blk + 512 * i,
                                       in ? rowp[i] + 512 * t + 128 * blk : rowp[i], in ? rbytes[i] : 0);
                    }
                    cp_async_arrive(&bars.fullB[s]);
                    TC2_TRACE(7, it);
                }
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (warp == 3 || warp == MMA_B_WARP) {
        // ===================== MMA issuers: warp 3 the forward units, warp 23 the backward tiles.  Two
        // in-order queues with blocking waits; a backward tile that is not ready (dPre, accumulator
        // not drained) never holds up the forward stream, which is what keeps HBM busy.  Each warp's
        // tcgen05.commit covers the MMAs its own elected lane issued.
        constexpr uint32_t TF32 = (1u << 4) | (2u << 7) | (2u << 10);
        if (warp == 3) {
            const uint32_t idesc_f = TF32 | (1u << 15) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // A MN-major, B K-major
            const uint64_t ad0 = make_desc(sbase + OFF_F + F_W, 4096, 512, 1);
            const uint64_t bd0 = make_desc(sbase + OFF_F + F_X, 16, 1024, 2);
            uint32_t itF = 0, itP = 0;
            for (int it_e = blockIdx.x; it_e < n_items; it_e += gridDim.x) {
                for (int u = 0; u < UF; ++u, ++itF) {
                    const int s = itF % SF, b = itP % NACC;
                    // The tensor core's adder truncates: a sum over all 98 K steps in ONE accumulator drifts by
                    // ~1e-6 of the result (and 25 steps still by ~3x the error of an fp32 FFMA chain, which the
                    // full-size parity test catches in the gradient ratios).  GRP units (8 MMAs) are summed per
                    // accumulator; the tail group reads it out and keeps the running sum in fp32 registers.
                    if (u % GRP == 0) mbar_wait(&bars.acc_free[b], ((itP / NACC) & 1) ^ 1, 7, dbg);
                    mbar_wait(&bars.convF[s], (itF / SF) & 1, 8, dbg);
                    tc_fence_after();
                    const bool last = u % GRP == GRP - 1 || u == UF - 1;
                    if (elect_one()) {
                        const uint64_t ad = ad0 + (uint64_t)(s * (F_STAGE >> 4)), bd = bd0 + (uint64_t)(s * (F_STAGE >> 4));
                        const uint32_t acc = tmem + TM_F + 64 * b;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)   // 8 features per MMA: 1024 bytes of A, 32 bytes of B
                            mma_tf32(acc, ad + (uint64_t)(kk * 64), bd + (uint64_t)(kk * 2), idesc_f, (u % GRP || kk) ? 1u : 0u);
                        mma_commit(&bars.emptyF[s]);
                        if (last) mma_commit(&bars.acc_full[b]);
                    }
                    __syncwarp();
                    if (last) ++itP;
                    TC2_TRACE(3, itF);
                }
            }
        } else {
            const uint32_t idesc_b = TF32 | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // both MN-major
            const uint64_t dpd = make_desc(sbase + OFF_DP, 4096, 512, 1);
            const uint64_t ahi0 = make_desc(sbase + OFF_B, 4096, 512, 1), alo0 = make_desc(sbase + OFF_B + B_XLO, 4096, 512, 1);
            uint32_t itB = 0;
            int k = 0;
            for (int it_e = blockIdx.x; it_e < n_items; it_e += gridDim.x, ++k) {
            const int e = env_at(it_e);
                mbar_wait(&bars.dpre_ready, k & 1, 4, dbg);
                for (int t = 0; t < TB; ++t, ++itB) {
                    const int s = itB % SB, g = itB & 1;
                    mbar_wait(&bars.convB[s], (itB / SB) & 1, 5, dbg);
                    mbar_wait(&bars.g_free[g], ((itB >> 1) & 1) ^ 1, 6, dbg);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ahi = ahi0 + (uint64_t)(s * (B_STAGE >> 4)), alo = alo0 + (uint64_t)(s * (B_STAGE >> 4));
                        const uint32_t acc = tmem + TM_G + 128 * g;
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) { // 8 samples per MMA = 1024 bytes of each block
                            mma_tf32(acc, ahi + (uint64_t)(kk * 64), dpd + (uint64_t)(kk * 64), idesc_b, kk ? 1u : 0u);
                            mma_tf32(acc, alo + (uint64_t)(kk * 64), dpd + (uint64_t)(kk * 64), idesc_b, 1u);
                        }
                        mma_commit(&bars.emptyB[s]);
                        mma_commit(&bars.g_full[g]);
                        if (t == TB - 1) mma_commit(&bars.dpre_free);
                    }
                    __syncwarp();
                    TC2_TRACE(4, itB);
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ===================== tail group (128 threads, named barrier 1)
        const int q = warp - 4, ttid = tid - 128;
        // scratch: HT [64 j][33] hidden activations (transposed, odd stride: lanes along j or along s
        // are both conflict free); tw = b1 | W2 padded to 16 classes per row | b2; tg = tail gradient
        // in parameter order; Zp [4][16][32] partial logits; dZ [32][16]
        float *HT = ts + T_H, *tw = ts + T_TW, *tg = ts + T_TG, *Zp = ts + T_ZP, *dZ = ts + T_Z, *gp = ts + T_GP;
        int *ys = reinterpret_cast<int *>(ts + T_YS);
        float *tmisc = ts + T_MISC;
        const int tailP = d.tailP;
        const float *b1 = tw, *W2p = tw + N1, *b2 = tw + N1 + N1 * CMAX;
        unsigned char *dP = sm + OFF_DP;
        constexpr int C4 = CC ? (CC + 3) / 4 : CMAX / 4;  // float4 groups of a padded class row that hold data
        for (int i = ttid; i < N1 * CMAX; i += 128)       // the padding of the W2 rows multiplies zeros of dZ: keep it finite
            if (i % CMAX >= C) tw[N1 + i] = 0.f;
        int k = 0;
        for (int it_e = blockIdx.x; it_e < n_items; it_e += gridDim.x, ++k) {
            const int e = env_at(it_e);
            const float *We = d.w + (size_t)e * d.Pp;
            float *gout = d.gnext + (size_t)e * d.Pp;
            const int *idx; int cnt;
            current_batch(d, a, e, d.sc + e, idx, cnt);
            for (int i = ttid; i < tailP; i += 128) {     // b1, W2 (rows padded to 16), b2
                const float v = We[d.P1 + i];
                const int r = i - N1;
                if (i < N1) tw[i] = v;
                else if (r < N1 * C) tw[N1 + (r / C) * CMAX + (r % C)] = v;
                else tw[N1 + N1 * CMAX + (r - N1 * C)] = v;
            }
            if (ttid < B) ys[ttid] = ttid < cnt ? d.labels[idx[ttid]] : 0;
            // ---- Hpre = X . W1 arrives from the read-out group in HT; + bias, relu (rows the minibatch does not have stay 0)
            mbar_wait(&bars.hpre_ready, k & 1, 9, dbg);
            if (q == 0) TC2_TRACE(5, 2 * k);
            {
                const int jj = ttid & 63, sg = ttid >> 6;
                const float bj = b1[jj];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int s = sg * 16 + i;
                    const float v = HT[jj * HS + s] + bj;
                    HT[jj * HS + s] = (s < cnt && v > 0.f) ? v : 0.f;
                }
            }
            group_bar(1);
            // ---- second layer (optimize_nn.py:42-44): lane = sample, warp = a quarter of the hidden units;
            // the W2 rows are broadcast reads, 16 x C FMAs on C independent accumulators per thread
            {
                float z[4 * C4];
#pragma unroll
                for (int c = 0; c < 4 * C4; ++c) z[c] = 0.f;
#pragma unroll 4
                for (int jj = 0; jj < 16; ++jj) {
                    const int jr = 16 * q + jj;
                    const float h = HT[jr * HS + lane];
#pragma unroll
                    for (int c4 = 0; c4 < C4; ++c4) {
                        const float4 w = *reinterpret_cast<const float4 *>(W2p + jr * CMAX + 4 * c4);
                        z[4 * c4] = fmaf(h, w.x, z[4 * c4]); z[4 * c4 + 1] = fmaf(h, w.y, z[4 * c4 + 1]);
                        z[4 * c4 + 2] = fmaf(h, w.z, z[4 * c4 + 2]); z[4 * c4 + 3] = fmaf(h, w.w, z[4 * c4 + 3]);
                    }
                }
#pragma unroll
                for (int c = 0; c < 4 * C4; ++c) Zp[(q * CMAX + c) * B + lane] = z[c];
            }
            group_bar(1);
            // ---- softmax cross-entropy and its derivative (optimize_nn.py:47-50): warp 0, lane = sample
            if (q == 0) {
                float z[4 * C4];
#pragma unroll
                for (int c = 0; c < 4 * C4; ++c)
                    z[c] = ((Zp[c * B + lane] + Zp[(CMAX + c) * B + lane]) + (Zp[(2 * CMAX + c) * B + lane] + Zp[(3 * CMAX + c) * B + lane])) +
                           (c < C ? b2[c] : 0.f);
                float loss = 0.f;
                if (lane < cnt) {
                    const int y = ys[lane];
                    float m = z[0];
#pragma unroll
                    for (int c = 1; c < 4 * C4; ++c) if (c < C) m = fmaxf(m, z[c]);
                    float sum = 0.f, zy = 0.f;
#pragma unroll
                    for (int c = 0; c < 4 * C4; ++c) if (c < C) { if (c == y) zy = z[c]; z[c] = expf(z[c] - m); sum += z[c]; }
                    loss = (m + logf(sum)) - zy;
                    const float inv = 1.0f / sum;
#pragma unroll
                    for (int c = 0; c < 4 * C4; ++c) z[c] = c < C ? z[c] * inv - (c == y ? 1.f : 0.f) : 0.f;
                } else {
#pragma unroll
                    for (int c = 0; c < 4 * C4; ++c) z[c] = 0.f;
                }
#pragma unroll
                for (int c4 = 0; c4 < C4; ++c4)
                    *reinterpret_cast<float4 *>(dZ + lane * CMAX + 4 * c4) = make_float4(z[4 * c4], z[4 * c4 + 1], z[4 * c4 + 2], z[4 * c4 + 3]);
                // loss = mean over the minibatch; gb2 = column sums of dZ: lanes are the samples
                float lsum = loss;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
                if (lane == 0) tmisc[0] = lsum / (float)cnt;
#pragma unroll
                for (int c = 0; c < 4 * C4; ++c) {
                    float g = z[c];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
                    if (lane == 0 && c < C) tg[N1 + N1 * C + c] = g;
                }
            }
            group_bar(1);
            // ---- backward of the tail.  gW2[j][c] = sum_s H[s][j] dZ[s][c]: thread = hidden unit j x half of the classes
            {
                const int jj = ttid & 63, half = ttid >> 6;
                constexpr int G0 = (C4 + 1) / 2;          // float4 class groups of the first half; the second has C4 - G0
                const int g_first = half * G0, g_count = half ? C4 - G0 : G0;
                float g[4 * G0];
#pragma unroll
                for (int c = 0; c < 4 * G0; ++c) g[c] = 0.f;
                for (int s = 0; s < cnt; ++s) {
                    const float h = HT[jj * HS + s];
#pragma unroll
                    for (int c4 = 0; c4 < G0; ++c4) {
                        if (c4 < g_count) {
                            const float4 z = *reinterpret_cast<const float4 *>(dZ + s * CMAX + 4 * (g_first + c4));
                            g[4 * c4] = fmaf(h, z.x, g[4 * c4]); g[4 * c4 + 1] = fmaf(h, z.y, g[4 * c4 + 1]);
                            g[4 * c4 + 2] = fmaf(h, z.z, g[4 * c4 + 2]); g[4 * c4 + 3] = fmaf(h, z.w, g[4 * c4 + 3]);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 4 * G0; ++c) {
                    const int cls = 4 * g_first + c;
                    if (c < 4 * g_count && cls < C) tg[N1 + jj * C + cls] = g[c];
                }
            }
            if (k >= 1) mbar_wait(&bars.dpre_free, (k - 1) & 1, 10, dbg);   // backward MMAs of the previous env have read dPre
            {   // dPre[s][j] = relu'(H[s][j]) sum_c dZ[s][c] W2[j][c]: thread = hidden unit j x 16 samples; W2 row in registers
                const int jj = ttid & 63, sg = ttid >> 6;
                float w2r[4 * C4];
#pragma unroll
                for (int c4 = 0; c4 < C4; ++c4) {
                    const float4 w = *reinterpret_cast<const float4 *>(W2p + jj * CMAX + 4 * c4);
                    w2r[4 * c4] = w.x; w2r[4 * c4 + 1] = w.y; w2r[4 * c4 + 2] = w.z; w2r[4 * c4 + 3] = w.w;
                }
                float colsum = 0.f;
                const uint32_t cbase = (uint32_t)((jj >> 5) * 4096 + ((jj & 7) << 2));
                const int q32 = (jj & 31) >> 3;
#pragma unroll 4
                for (int i = 0; i < 16; ++i) {
                    const int s = sg * 16 + i;
                    float v = 0.f;
#pragma unroll
                    for (int c4 = 0; c4 < C4; ++c4) {     // padded classes hold zeros in dZ and W2p
                        const float4 z = *reinterpret_cast<const float4 *>(dZ + s * CMAX + 4 * c4);
                        v = fmaf(z.x, w2r[4 * c4], v); v = fmaf(z.y, w2r[4 * c4 + 1], v);
                        v = fmaf(z.z, w2r[4 * c4 + 2], v); v = fmaf(z.w, w2r[4 * c4 + 3], v);
                    }
                    v = HT[jj * HS + s] > 0.f ? v : 0.f;  // rows beyond the minibatch have H = 0
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
            if (q == 0) TC2_TRACE(5, 2 * k + 1);
            if (ttid < N1) tg[ttid] = gp[ttid] + gp[N1 + ttid];          // gb1
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
            if (ttid == 0) mbar_arrive(&bars.hpre_free);  // the read-out group may write the next env's Hpre
        }
    } else if (warp >= 8 && warp < 12) {
        // ===================== drain group (128 threads, named barrier 2): gradient tiles -> HBM
        const int q = warp - 8, dtid = tid - 256;
        unsigned char *stg = sm + OFF_G + q * 4096;
        uint32_t itG = 0;
        int k = 0;
        for (int it_e = blockIdx.x; it_e < n_items; it_e += gridDim.x, ++k) {
            const int e = env_at(it_e);
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
#pragma unroll
                    for (int half = 0; half < 2; ++half) {            // 16 columns at a time keeps the warp under 80 registers
                        uint32_t v0[16], v1[16];
                        TC2_LD16(taddr + 64 + 32 * h + 16 * half, v1);        // x . dPre_lo
                        TC2_LD16(taddr + 32 * h + 16 * half, v0);             // x . dPre_hi
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[16 * half + i] = __uint_as_float(v1[i]) + __uint_as_float(v0[i]);
                    }
                    if (h == 1) {                                     // accumulator drained: the next tile may overwrite it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bars.g_free[g]);
                        if (q == 0) TC2_TRACE(6, itG);
                    }
                    // lane = feature row: the 32 x 32 block goes through the warp's staging block (128-byte
                    // rows, SWIZZLE_128B) and leaves as ONE bulk tensor store; rows beyond D are clipped
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous block has left the staging block
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        *reinterpret_cast<float4 *>(stg + lane * 128 + swz16(c, lane)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                        gsum += (v[4 * c] + v[4 * c + 1]) + (v[4 * c + 2] + v[4 * c + 3]);
                    }
                    fence_async();
                    __syncwarp();
                    if (lane == 0 && f0 < D) {
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                                     ::"l"(gsel ? &map_g1 : &map_g0), "r"(32 * h), "r"(f0), "r"(e), "r"(smem_u32(stg)) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
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
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the last gradient blocks have reached memory
    } else {
        // ===================== converters: lo parts of the landed tiles.  Warps 12-19 serve the forward
        // ring, warps 20-22 the backward ring, so that neither queue waits for the other.  Many
        // threads with a few float4 each: what counts is the latency of a unit, not the throughput
        // (every microsecond a stage spends here is a microsecond it is not in flight to HBM).
        if (warp >= 12 && warp < 12 + CONV_F_WARPS) {
            const int cid = tid - 384;                    // 0..127
            uint32_t itF = 0;
            for (int it_e = blockIdx.x; it_e < n_items; it_e += gridDim.x) {
                for (int u = 0; u < UF; ++u, ++itF) {
                    const int s = itF % SF;
                    mbar_wait(&bars.fullF[s], (itF / SF) & 1, 14, dbg);
                    if (warp == 12) TC2_TRACE(1, itF);
                    unsigned char *base = sm + OFF_F + s * F_STAGE;
                    // loads first, stores after: the compiler cannot move a shared-memory load above an
                    // earlier store to the same array
                    float4 v[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) {         // 512 float4 of W, then 256 of X
                        const int c = i * 128 + cid;
                        v[i] = *reinterpret_cast<const float4 *>(base + (i < 4 ? F_W + c * 16 : F_X + (c - 512) * 16));
                    }
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        const int c = i * 128 + cid;
                        *reinterpret_cast<float4 *>(base + (i < 4 ? F_WLO + c * 16 : F_XLO + (c - 512) * 16)) = lo_of4(v[i]);
                    }
                    fence_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.convF[s]);
                    if (warp == 12) TC2_TRACE(2, itF);
                }
            }
        } else if (warp >= READ_WARP0 && warp < READ_WARP0 + 4) {
            // ===================== read-out group (128 threads, named barrier 3): the forward accumulators.
            // Lanes 0..63 of an accumulator = W_hi rows, 64..127 = W_lo rows; columns 0..31 = X_hi, 32..63 = X_lo.
            // One read-out per GRP units; the running sum of the env lives in registers (fp32, round to
            // nearest); at the end of the env the four partial products meet in HT [64 j][32 s].
            const int q = warp - READ_WARP0;
            float *HT = ts + T_H;
            uint32_t itP = 0;
            int k = 0;
            for (int it_e = blockIdx.x; it_e < n_items; it_e += gridDim.x, ++k) {
            const int e = env_at(it_e);
                float acc[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = 0.f;
                for (int u0 = 0; u0 < UF; u0 += GRP, ++itP) {
                    const int b = itP % NACC;
                    mbar_wait(&bars.acc_full[b], (itP / NACC) & 1, 15, dbg);
                    tc_fence_after();
                    const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + TM_F + 64 * b;
#pragma unroll
                    for (int half = 1; half >= 0; --half) {   // the X_lo columns (small terms) first
                        uint32_t v[32];
                        TC2_LD32(taddr + 32 * half, v);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < 32; ++i) acc[i] += __uint_as_float(v[i]);
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.acc_free[b]);
                }
                if (k >= 1) mbar_wait(&bars.hpre_free, (k - 1) & 1, 16, dbg);     // the tail is done with the previous env's activations
                const int j = (q & 1) * 32 + lane;
                if (q >= 2) {
#pragma unroll
                    for (int s = 0; s < B; ++s) HT[j * HS + s] = acc[s];
                }
                asm volatile("bar.sync 3, 128;" ::: "memory");
                if (q < 2) {
#pragma unroll
                    for (int s = 0; s < B; ++s) HT[j * HS + s] += acc[s];
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.hpre_ready);
                }
            }
        } else if (warp >= CONV_B_WARP0 && warp < CONV_B_WARP0 + CONV_B_WARPS) {
            const int cid = tid - 32 * CONV_B_WARP0;        // 0..95
            uint32_t itB = 0;
            for (int it_e = blockIdx.x; it_e < n_items; it_e += gridDim.x) {
                for (int t = 0; t < TB; ++t, ++itB) {
                    const int s = itB % SB;
                    mbar_wait(&bars.fullB[s], (itB / SB) & 1, 13, dbg);
                    unsigned char *base = sm + OFF_B + s * B_STAGE;
                    constexpr int CT = 32 * CONV_B_WARPS, PER = (1024 + CT - 1) / CT;     // 96 threads, 11 float4 each
                    float4 v[PER];
#pragma unroll
                    for (int i = 0; i < PER; ++i)
                        if (i * CT + cid < 1024) v[i] = *reinterpret_cast<const float4 *>(base + (i * CT + cid) * 16);
#pragma unroll
                    for (int i = 0; i < PER; ++i)
                        if (i * CT + cid < 1024) *reinterpret_cast<float4 *>(base + B_XLO + (i * CT + cid) * 16) = lo_of4(v[i]);
                    fence_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars.convB[s]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace tc2
}  // namespace

// ---------------------------------------------------------------- host side
struct b2e_tc2_ctx {
    CUtensorMap map_w;               // parameters as [E][2 column halves][D][32] fp32, box {32, 32, 2, 1}
    CUtensorMap map_g[2];            // the two gradient buffers (they swap roles every step) as [E][D][64], box {32, 32, 1}
    const float *g_base[2];
    int pf_units;                    // B2E_TC_PF=0 switches the L2 prefetch of the next env's parameters off
    int *dbg;                        // 16 watchdog ints, then 8 x 512 trace slots (long long)
    int grid;
    std::string trace_path;
};

bool b2e_tc2_supported(const void *dev) {
    const Dev &d = *static_cast<const Dev *>(dev);
    return d.kind == B2E_PROBLEM_SOFTMAX && d.hidden && !d.generic && d.N1 == tc2::N1 && d.B == tc2::B &&
           d.C >= 1 && d.C <= tc2::CMAX && d.D >= 32 && d.D % 4 == 0 && d.Dp == d.D && d.Pp % 4 == 0;
}

namespace {
template <bool SECOND, int CC>
bool tc2_set_smem() {
    return cudaFuncSetAttribute(tc2::tc2_eval_kernel<SECOND, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES) == cudaSuccess;
}
}  // namespace

b2e_tc2_ctx *b2e_tc2_create(const void *dev, int num_sms, std::string *error) {
    const Dev &d = *static_cast<const Dev *>(dev);
    if (!(tc2_set_smem<false, 10>() && tc2_set_smem<true, 10>() && tc2_set_smem<false, 0>() && tc2_set_smem<true, 0>())) {
        *error = "tcgen05 eval kernel does not fit shared memory";
        return nullptr;
    }
    b2e_tc2_ctx *ctx = new (std::nothrow) b2e_tc2_ctx();
    if (!ctx) { *error = "out of host memory"; return nullptr; }
    ctx->grid = d.E < num_sms ? d.E : num_sms;
    ctx->dbg = nullptr;
    if (cudaMalloc((void **)&ctx->dbg, tc2::DBG_BYTES) != cudaSuccess || cudaMemset(ctx->dbg, 0, tc2::DBG_BYTES) != cudaSuccess) {
        *error = "cudaMalloc of the watchdog word failed";
        delete ctx;
        return nullptr;
    }
    if (const char *tp = getenv("B2E_TC_TRACE")) {
        ctx->trace_path = tp;
        const int one = 1;
        cudaMemcpy(ctx->dbg + 15, &one, sizeof(int), cudaMemcpyHostToDevice);
    }
    // gradient blocks leave as bulk tensor stores: [E][D][64] fp32 tensors over the two gradient buffers, box {32, 32, 1}
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || !fn) {
        *error = "cuTensorMapEncodeTiled is not available in this driver";
        b2e_tc2_destroy(ctx);
        return nullptr;
    }
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    const cuuint64_t dims[3] = {(cuuint64_t)tc2::N1, (cuuint64_t)d.D, (cuuint64_t)d.E};
    const cuuint64_t strides[2] = {(cuuint64_t)tc2::N1 * 4, (cuuint64_t)d.Pp * 4};
    const cuuint32_t box[3] = {32, 32, 1}, estr[4] = {1, 1, 1, 1};
    auto encode = [&](CUtensorMap *map, const float *base, CUtensorMapSwizzle swizzle) {
        return reinterpret_cast<encode_fn>(fn)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, box, estr,
                                               CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    };
    ctx->g_base[0] = d.gnext;
    ctx->g_base[1] = d.gprev;
    const cuuint64_t wdims[4] = {32, (cuuint64_t)d.D, 2, (cuuint64_t)d.E};
    const cuuint64_t wstrides[3] = {(cuuint64_t)tc2::N1 * 4, 128, (cuuint64_t)d.Pp * 4};
    const cuuint32_t wbox[4] = {32, 32, 2, 1};
    if (reinterpret_cast<encode_fn>(fn)(&ctx->map_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)d.w, wdims, wstrides, wbox, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
        !encode(&ctx->map_g[0], d.gnext, CU_TENSOR_MAP_SWIZZLE_128B) ||
        !encode(&ctx->map_g[1], d.gprev, CU_TENSOR_MAP_SWIZZLE_128B)) {
        *error = "cuTensorMapEncodeTiled failed";
        b2e_tc2_destroy(ctx);
        return nullptr;
    }
    return ctx;
}

void b2e_tc2_destroy(b2e_tc2_ctx *ctx) {
    if (!ctx) return;
    if (!ctx->trace_path.empty()) {                      // timeline of CTA 0 in the last launch
        static long long host[8 * 512];
        if (cudaDeviceSynchronize() == cudaSuccess &&
            cudaMemcpy(host, ctx->dbg + 16, sizeof(host), cudaMemcpyDeviceToHost) == cudaSuccess) {
            if (FILE *fh = fopen(ctx->trace_path.c_str(), "w")) {
                const char *names[8] = {"w_issue", "w_landed", "conv_f", "mma_f", "mma_b", "tail", "drain", "xb_issue"};
                for (int r = 0; r < 8; ++r) {
                    fprintf(fh, "%s", names[r]);
                    for (int i = 0; i < 512 && host[r * 512 + i]; ++i) fprintf(fh, " %lld", host[r * 512 + i]);
                    fprintf(fh, "\n");
                }
                fclose(fh);
            }
        }
    }
    cudaFree(ctx->dbg);
    delete ctx;
}

int b2e_tc2_grid(const b2e_tc2_ctx *ctx) { return ctx->grid; }

int b2e_tc2_watchdog(const b2e_tc2_ctx *ctx, int out[4]) {
    return cudaMemcpy(out, ctx->dbg, 4 * sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 1;
}

int b2e_tc2_launch(const b2e_tc2_ctx *ctx, const void *dev, const void *args, int second, void *stream) {
    const Dev &d = *static_cast<const Dev *>(dev);
    const StepArgs &a = *static_cast<const StepArgs *>(args);
    const cudaStream_t cs = (cudaStream_t)stream;
    const int grid = ctx->grid;
    const int gsel = d.gnext == ctx->g_base[0] ? 0 : 1;
    if (d.gnext != ctx->g_base[gsel]) return 1;          // not one of the two gradient buffers
    const CUtensorMap &m0 = ctx->map_g[0], &m1 = ctx->map_g[1];
    if (d.C == 10) {
        if (second) tc2::tc2_eval_kernel<true, 10><<<grid, tc2::THREADS, tc2::SMEM_BYTES, cs>>>(d, a, ctx->map_w, m0, m1, gsel, ctx->pf_units, ctx->dbg);
        else tc2::tc2_eval_kernel<false, 10><<<grid, tc2::THREADS, tc2::SMEM_BYTES, cs>>>(d, a, ctx->map_w, m0, m1, gsel, ctx->pf_units, ctx->dbg);
    } else {
        if (second) tc2::tc2_eval_kernel<true, 0><<<grid, tc2::THREADS, tc2::SMEM_BYTES, cs>>>(d, a, ctx->map_w, m0, m1, gsel, ctx->pf_units, ctx->dbg);
        else tc2::tc2_eval_kernel<false, 0><<<grid, tc2::THREADS, tc2::SMEM_BYTES, cs>>>(d, a, ctx->map_w, m0, m1, gsel, ctx->pf_units, ctx->dbg);
    }
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

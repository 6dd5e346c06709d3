// Shared per-agent policy of libb200env (C ABI: include/b200policy.h).
//
// What the reference's callers do with the observation matrix of OptVecEnv.step_wait
// (vectorize/optvecenv.py:78-88): stable-baselines' runner evaluates ONE MlpPolicy -- two tanh layers
// of 64 units and a linear head -- on each of the sum(P) agent rows (run_multiagent_exp_single.py:37-49,
// play_optimize.py:79-98).  At BASELINE config 4 that is 2.08e8 rows x (15*64 + 64*64 + 64) MACs =
// 2.1 TFLOP per env step: dense GEMM work, so it runs on the tensor cores:
//
//   tile = 128 agent rows; A1 [128 x 16] bf16 = the observation rows + a column of ones (bias),
//   D1 [128 x 64] = A1 . [W1 | b1]^T         one tcgen05.mma kind::f16 (M 128, N 64, K 16), fp32 in TMEM
//   A2 [128 x 80] bf16 = tanh(D1) + a column of ones + zero padding
//   D2 [128 x 64] = A2 . [W2 | b2 | 0]^T     five MMAs
//   mean = tanh(D2) . w3 + b3                 FFMA on the fp32 accumulators, lane = row
//
// Persistent, one 768-thread CTA per SM, SIX independent groups of four warps; every group runs its own
// tiles start to finish (operand staging, MMA issue by its thread 0, tcgen05.commit -> the group's
// mbarrier, tcgen05.ld epilogues), so while one group waits for its MMA the SM's MUFU units -- the
// bound of this kernel: 128 tanh per row, 16 per clock per SM -- are busy with the other five (four groups
// reach 0.63 of that rate, six 0.75).  TMEM: 64 columns per group, D2 overwrites D1 (read out before).  Operands use the no-swizzle K-major canonical layout: core
// matrices of 8 rows x 16 bytes, contiguous (128 B), LBO = 128 (next core matrix along K), SBO = K/8 * 128
// (next 8 rows).
//
// Two front ends: dense rows (what b2e_step wrote; one 7.5 KB bulk copy per tile, requested a tile ahead) and
// the env's adjusted-history rings (MultiOptLRs; lane = parameter, 2H coalesced 4-byte loads prefetched one
// tile ahead; the 3H observation words per agent never exist in HBM).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <new>
#include <string>

#include "b200env_shared.cuh"
#include "b200env_internal.h"
#include "b200policy.h"

namespace {
namespace pol {

#ifndef B2P_GROUPS
#define B2P_GROUPS 6
#endif
constexpr int TILE = 128, HID = B2P_HIDDEN, K1 = 16, K2 = 80, GROUPS = B2P_GROUPS, THREADS = GROUPS * 128;
constexpr int NSTG = GROUPS <= 4 ? 2 : 1;        // staging buffers of the dense front end per group (shared memory budget)
constexpr int XMAX = B2P_MAX_OBS_DIM;
constexpr int SBO1 = (K1 / 8) * 128, SBO2 = (K2 / 8) * 128;
constexpr int A1_BYTES = (TILE / 8) * SBO1, A2_BYTES = (TILE / 8) * SBO2;
constexpr int STAGE_BYTES = TILE * XMAX * 4;
constexpr int G_BYTES = NSTG * STAGE_BYTES + A1_BYTES + A2_BYTES;
constexpr int OFF_B1 = GROUPS * G_BYTES, B1_BYTES = (HID / 8) * SBO1;
constexpr int OFF_B2 = OFF_B1 + B1_BYTES, B2_BYTES = (HID / 8) * SBO2;
constexpr int OFF_W3 = OFF_B2 + B2_BYTES;
constexpr int SMEM_BYTES = OFF_W3 + (HID + 4) * 4 + 128;   // + slack to align the base to 128 bytes
constexpr int TMEM_COLS = GROUPS <= 4 ? 256 : 512;    // 64 columns per group: D2 overwrites D1, which is fully read before the second layer's MMAs are issued
static_assert(G_BYTES % 128 == 0 && OFF_B1 % 128 == 0 && OFF_B2 % 128 == 0, "operand blocks are 128-byte aligned");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");

struct Args {
    const float *w1, *b1, *w2, *b2, *w3, *b3;
    const float *obs;                 // dense front end: [rows, obs_dim]
    float *out;
    long long rows, tiles;
    int units, segs;                  // ring front end: units = envs x segments of an env's tiles
    unsigned long long seed;
    const unsigned long long *seed_counter;   // added to seed when set (b2p_set_seed_counter)
    float noise_std, low, high;
    int obs_dim, bulk_ok;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

#define B2P_LD32(taddr, v) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
                 : "r"(taddr))

#define B2P_LD16(taddr, v) \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
                 : "r"(taddr))

__device__ __forceinline__ float tanh_f32(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t tanh_bf16x2(uint32_t x) {
    uint32_t y;
    asm("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
// {lo at the lower address, hi above it}, round to nearest even
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ unsigned short bf16_bits(float v) { return (unsigned short)(pack_bf16x2(v, 0.f) & 0xFFFFu); }

// N(0, 1) of an agent row: counter-based (splitmix64 of seed and row), Box-Muller
__device__ __forceinline__ float row_noise(unsigned long long seed, long long row) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(row + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u1 = ((float)(unsigned)(z >> 40) + 1.0f) * (1.0f / 16777216.0f);       // (0, 1]
    const float u2 = (float)(unsigned)(z & 0xFFFFFFu) * (1.0f / 16777216.0f);          // [0, 1)
    return sqrtf(-2.0f * __logf(u1)) * cospif(2.0f * u2);
}

// FRONT: 0 = dense observation rows, 1 = rings with max_history 5 (every index a compile-time constant),
// 2 = rings with any max_history <= 5
template <int FRONT, int TANH>
__global__ void __launch_bounds__(THREADS, 1) policy_kernel(const __grid_constant__ Args a, const __grid_constant__ Dev d) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t obs_full[GROUPS][2];
    __shared__ __align__(8) uint64_t mma_bar[GROUPS];
    __shared__ uint32_t tmem_slot;
    constexpr bool RING = FRONT != 0;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int g = warp >> 2, wq = warp & 3, gt = tid & 127;          // group, TMEM lane quarter, row of the tile
    const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
    unsigned char *const sm = smem_raw + (sbase - smem_u32(smem_raw));
    unsigned char *const gsm = sm + g * G_BYTES;
    unsigned char *const A1 = gsm + NSTG * STAGE_BYTES, *const A2 = A1 + A1_BYTES;
    float *const w3s = reinterpret_cast<float *>(sm + OFF_W3);
    const int od = a.obs_dim;

    if (tid == 0) {
        for (int i = 0; i < GROUPS; ++i) { mbar_init(&obs_full[i][0], 1); mbar_init(&obs_full[i][1], 1); mbar_init(&mma_bar[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // the network as bf16 B operands [N = 64 units][K], biases as one more K column (the A operands carry a 1 there)
    for (int i = tid; i < HID * K1; i += THREADS) {
        const int n = i / K1, k = i - n * K1;
        const float v = k < od ? a.w1[n * od + k] : (k == K1 - 1 ? a.b1[n] : 0.f);
        *reinterpret_cast<unsigned short *>(sm + OFF_B1 + (n >> 3) * SBO1 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) = bf16_bits(v);
    }
    for (int i = tid; i < HID * K2; i += THREADS) {
        const int n = i / K2, k = i - n * K2;
        const float v = k < HID ? a.w2[n * HID + k] : (k == HID ? a.b2[n] : 0.f);
        *reinterpret_cast<unsigned short *>(sm + OFF_B2 + (n >> 3) * SBO2 + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) = bf16_bits(v);
    }
    if (tid < HID) w3s[tid] = a.w3[tid];
    if (tid == HID) w3s[HID] = a.b3[0];
    {   // constant columns of this thread's A2 row: k = 64 is the bias 1, k = 65..79 are zero
        unsigned char *row = A2 + (gt >> 3) * SBO2 + (gt & 7) * 16;
        *reinterpret_cast<uint4 *>(row + 8 * 128) = make_uint4(0x00003F80u, 0u, 0u, 0u);
        *reinterpret_cast<uint4 *>(row + 9 * 128) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HID >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);   // f32 += bf16 . bf16, both K-major
    const uint64_t a1d = make_desc(sbase + g * G_BYTES + NSTG * STAGE_BYTES, 128, SBO1);
    const uint64_t a2d = make_desc(sbase + g * G_BYTES + NSTG * STAGE_BYTES + A1_BYTES, 128, SBO2);
    const uint64_t b1d = make_desc(sbase + OFF_B1, 128, SBO1), b2d = make_desc(sbase + OFF_B2, 128, SBO2);
    const uint32_t tm1 = tmem + 64 * g, tm2 = tm1;
    const uint32_t tlane = (uint32_t)(wq * 32) << 16;
    const float b3 = w3s[HID];
    const unsigned long long seed = a.seed + (a.seed_counter ? *a.seed_counter : 0ull);

    // ---- front ends
    const int tiles_per_env = RING ? (d.P + TILE - 1) / TILE : 1;
    auto fetch_dense = [&](long long tile, int b) {         // observation rows of `tile` -> stage b, completes on obs_full[g][b]
        const long long row0 = tile * TILE;
        const int nrows = (int)min((long long)TILE, a.rows - row0);
        const unsigned bytes = (unsigned)(nrows * od) * 4u;
        const float *src = a.obs + row0 * od;
        float *dst = reinterpret_cast<float *>(gsm + b * STAGE_BYTES);
        if (a.bulk_ok && (bytes & 15u) == 0u) {
            if (gt == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&obs_full[g][b])), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(&obs_full[g][b])) : "memory");
            }
        } else {                                            // ragged last tile / unaligned matrix: plain loads
            for (int i = gt; i < nrows * od; i += 128) dst[i] = src[i];
            group_bar(g);
            if (gt == 0) mbar_arrive(&obs_full[g][b]);
        }
    };
    // Ring front end.  A CTA owns "units" (env, segment of the env's tiles), u = blockIdx.x + k * gridDim.x,
    // and its four groups share a unit's tiles round robin: the env's ring position (head, nvalid) and loss
    // history sit behind a dependent load and are fetched once per unit, not once per tile.  The cursor is
    // always ONE TILE AHEAD of the tile being computed: load_ring issues the 2H loads of that tile into
    // registers (raw values) under the current tile's work, ring_row turns them into the observation row
    // when the tile's turn comes.
    struct Cursor {
        long long tile;               // dense: tile index
        int unit, tp, hi, e, head, nvalid;
        bool valid;
        float lv[XMAX / 3];           // adjusted losses newest first (0 beyond nvalid)
    };
    auto ring_enter = [&](Cursor &c) {                      // first tile of this group in unit c.unit or a later one
        for (; c.unit < a.units; c.unit += (int)gridDim.x) {
            const int e = c.unit / a.segs, sg = c.unit - e * a.segs;
            const int lo = (int)((long long)sg * tiles_per_env / a.segs), hi = (int)((long long)(sg + 1) * tiles_per_env / a.segs);
            if (lo + g >= hi) continue;
            const EnvScalars *sc = d.sc + e;
            c.e = e; c.tp = lo + g; c.hi = hi; c.head = sc->head; c.nvalid = sc->nvalid; c.valid = true;
#pragma unroll
            for (int h = 0; h < XMAX / 3; ++h) {
                const int H = FRONT == 1 ? 5 : d.H;
                int slot = c.head - h;
                slot += slot < 0 ? H : 0;
                c.lv[h] = (h < H && h < c.nvalid) ? sc->adj_loss[slot] : 0.f;
            }
            return;
        }
        c.valid = false;
    };
    auto advance = [&](Cursor &c) {
        if (RING) {
            c.tp += GROUPS;
            if (c.tp >= c.hi) { c.unit += (int)gridDim.x; ring_enter(c); }
        } else {
            c.tile += (long long)gridDim.x * GROUPS;
            c.valid = c.tile < a.tiles;
        }
    };
    auto load_ring = [&](const Cursor &c, float (&x)[XMAX]) {
        const int p = c.tp * TILE + gt, H = FRONT == 1 ? 5 : d.H;
        const bool ok = p < d.P;
#pragma unroll
        for (int h = 0; h < XMAX / 3; ++h) {
            float wv = 0.f, gv = 0.f;
            if (h < H) {
                int slot = c.head - h;
                slot += slot < 0 ? H : 0;
                const size_t off = ((size_t)c.e * H + slot) * d.Pp + p;
                if (ok && h < c.nvalid) { wv = d.ringw[off]; gv = d.ringg[off]; }
            }
            x[3 * h] = wv; x[3 * h + 1] = c.lv[h]; x[3 * h + 2] = gv;
        }
    };
    auto ring_row = [&](const float (&raw)[XMAX], float (&x)[XMAX]) {   // [adj_w (H) | adj_L (H) | adj_g (H)], clip, -1
#pragma unroll
        for (int k = 0; k < XMAX; ++k) x[k] = 0.f;
        if (FRONT == 1) {
#pragma unroll
            for (int h = 0; h < 5; ++h) { x[h] = clip_m1(raw[3 * h]); x[5 + h] = clip_m1(raw[3 * h + 1]); x[10 + h] = clip_m1(raw[3 * h + 2]); }
        } else {
            const int H = d.H;                               // rare shapes: the row is assembled with selects
#pragma unroll
            for (int k = 0; k < XMAX; ++k) {
                float v = 0.f;
#pragma unroll
                for (int h = 0; h < XMAX / 3; ++h) {
                    v = (k == h && h < H) ? raw[3 * h] : v;
                    v = (k == H + h && h < H) ? raw[3 * h + 1] : v;
                    v = (k == 2 * H + h && h < H) ? raw[3 * h + 2] : v;
                }
                x[k] = k < 3 * H ? clip_m1(v) : 0.f;
            }
        }
    };

    Cursor c;
    c.tile = (long long)blockIdx.x * GROUPS + g;
    c.unit = (int)blockIdx.x;
    c.valid = c.tile < a.tiles;
    if (RING) ring_enter(c);
    uint32_t mph = 0;
    float nxt[XMAX];
    long long tile = 0;               // the tile being computed: dense index / (env, tile of the env)
    int cur_e = 0, cur_tp = 0;
    bool have = c.valid;
    if (have) {
        if (RING) { load_ring(c, nxt); cur_e = c.e; cur_tp = c.tp; }
        else { fetch_dense(c.tile, 0); tile = c.tile; }
        advance(c);
    }
    for (int n = 0; have; ++n) {
        float x[XMAX];
        if (RING) {
            ring_row(nxt, x);
        } else {
            const int b = n % NSTG;
            if (NSTG == 2 && c.valid) fetch_dense(c.tile, b ^ 1);
            mbar_wait(&obs_full[g][b], (uint32_t)(n / NSTG) & 1u);
            const float *st = reinterpret_cast<const float *>(gsm + b * STAGE_BYTES) + gt * od;
            const bool live = tile * TILE + gt < a.rows;
#pragma unroll
            for (int k = 0; k < XMAX; ++k) x[k] = (live && k < od) ? st[k] : 0.f;
        }
        // ---- A1 = [x | 0 | 1] as bf16, this thread's row
        {
            unsigned char *row = A1 + (gt >> 3) * SBO1 + (gt & 7) * 16;
            *reinterpret_cast<uint4 *>(row) = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]),
                                                         pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
            *reinterpret_cast<uint4 *>(row + 128) = make_uint4(pack_bf16x2(x[8], x[9]), pack_bf16x2(x[10], x[11]),
                                                               pack_bf16x2(x[12], x[13]), pack_bf16x2(x[14], 1.0f));
        }
        fence_async();
        tc_fence_before();
        group_bar(g);
        if (gt == 0) {
            tc_fence_after();
            mma_bf16(tm1, a1d, b1d, idesc, 0u);
            mma_commit(&mma_bar[g]);
        }
        if (!RING && NSTG == 1 && c.valid) fetch_dense(c.tile, 0);      // the single staging buffer is free again (barrier above)
        const bool have_next = c.valid;
        const long long next_tile = c.tile;
        const int next_e = c.e, next_tp = c.tp;
        if (RING && have_next) load_ring(c, nxt);             // in flight under this tile's work
        if (have_next) advance(c);
        mbar_wait(&mma_bar[g], mph);
        mph ^= 1u;
        tc_fence_after();
        // ---- A2 = tanh(D1) as bf16, 16 accumulator columns at a time (register budget of 768 threads)
#pragma unroll
        for (int qt = 0; qt < 4; ++qt) {
            uint32_t v[16];
            B2P_LD16(tm1 + tlane + 16 * qt, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            unsigned char *row = A2 + (gt >> 3) * SBO2 + (gt & 7) * 16 + (2 * qt) * 128;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t pk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float v0 = __uint_as_float(v[8 * c + 2 * i]), v1 = __uint_as_float(v[8 * c + 2 * i + 1]);
                    pk[i] = TANH >= 1 ? tanh_bf16x2(pack_bf16x2(v0, v1)) : pack_bf16x2(tanh_f32(v0), tanh_f32(v1));
                }
                *reinterpret_cast<uint4 *>(row + c * 128) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
        fence_async();
        tc_fence_before();
        group_bar(g);
        if (gt == 0) {
            tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < K2 / 16; ++kk)              // 16 K elements = two core matrices = 256 bytes per MMA
                mma_bf16(tm2, a2d + (uint64_t)(kk * 16), b2d + (uint64_t)(kk * 16), idesc, kk ? 1u : 0u);
            mma_commit(&mma_bar[g]);
        }
        mbar_wait(&mma_bar[g], mph);
        mph ^= 1u;
        tc_fence_after();
        // ---- head: mean = tanh(D2) . w3 + b3
        float acc0 = b3, acc1 = 0.f;
#pragma unroll
        for (int qt = 0; qt < 4; ++qt) {
            uint32_t v[16];
            B2P_LD16(tm2 + tlane + 16 * qt, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 w = *reinterpret_cast<const float4 *>(w3s + 16 * qt + 4 * c);
                float t0, t1, t2, t3;
                if (TANH == 2) {
                    const uint32_t p0 = tanh_bf16x2(pack_bf16x2(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1])));
                    const uint32_t p1 = tanh_bf16x2(pack_bf16x2(__uint_as_float(v[4 * c + 2]), __uint_as_float(v[4 * c + 3])));
                    t0 = __uint_as_float(p0 << 16); t1 = __uint_as_float(p0 & 0xFFFF0000u);
                    t2 = __uint_as_float(p1 << 16); t3 = __uint_as_float(p1 & 0xFFFF0000u);
                } else {
                    t0 = tanh_f32(__uint_as_float(v[4 * c])); t1 = tanh_f32(__uint_as_float(v[4 * c + 1]));
                    t2 = tanh_f32(__uint_as_float(v[4 * c + 2])); t3 = tanh_f32(__uint_as_float(v[4 * c + 3]));
                }
                acc0 = fmaf(t0, w.x, acc0); acc1 = fmaf(t1, w.y, acc1);
                acc0 = fmaf(t2, w.z, acc0); acc1 = fmaf(t3, w.w, acc1);
            }
        }
        tc_fence_before();
        float mean = acc0 + acc1;
        // ---- action of the row (the next b2e_step gathers it through the same row table)
        long long orow;
        bool live;
        if (RING) {
            const int p = cur_tp * TILE + gt;
            live = p < d.P;
            orow = (long long)cur_e * d.P + (live ? (d.row_lex ? d.row_of_param[p] : p) : 0);
        } else {
            orow = tile * TILE + gt;
            live = orow < a.rows;
        }
        if (a.noise_std != 0.f) mean = fmaf(a.noise_std, row_noise(seed, orow), mean);
        if (live) a.out[orow] = fminf(fmaxf(mean, a.low), a.high);
        have = have_next; tile = next_tile; cur_e = next_e; cur_tp = next_tp;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

}  // namespace pol
}  // namespace

// ---------------------------------------------------------------- host side (C ABI)
struct b2p_policy {
    int device, obs_dim, tanh_mode, num_sms;
    float *weights;                  // w1 [64, obs_dim] | b1 | w2 [64, 64] | b2 | w3 | b3
    const unsigned long long *seed_counter;
    bool have_weights;
    std::string error;
};

namespace {

std::string g_policy_create_error;

int pfail(b2p_handle h, const std::string &msg) {
    if (h) h->error = msg; else g_policy_create_error = msg;
    return 1;
}

struct PolicyDeviceGuard {
    int prev;
    bool switched;
    explicit PolicyDeviceGuard(int device) : prev(-1), switched(false) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
    }
    ~PolicyDeviceGuard() { if (switched) cudaSetDevice(prev); }
};

template <int FRONT>
cudaError_t launch_policy(int tanh_mode, int grid, cudaStream_t cs, const pol::Args &a, const Dev &d) {
    switch (tanh_mode) {
        case 2: pol::policy_kernel<FRONT, 2><<<grid, pol::THREADS, pol::SMEM_BYTES, cs>>>(a, d); break;
        case 1: pol::policy_kernel<FRONT, 1><<<grid, pol::THREADS, pol::SMEM_BYTES, cs>>>(a, d); break;
        default: pol::policy_kernel<FRONT, 0><<<grid, pol::THREADS, pol::SMEM_BYTES, cs>>>(a, d); break;
    }
    return cudaGetLastError();
}

template <int FRONT>
bool set_policy_smem() {
    return cudaFuncSetAttribute(pol::policy_kernel<FRONT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, pol::SMEM_BYTES) == cudaSuccess &&
           cudaFuncSetAttribute(pol::policy_kernel<FRONT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, pol::SMEM_BYTES) == cudaSuccess &&
           cudaFuncSetAttribute(pol::policy_kernel<FRONT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, pol::SMEM_BYTES) == cudaSuccess;
}

void fill_weights(const b2p_policy *h, pol::Args &a) {
    const int od = h->obs_dim, hid = pol::HID;
    const float *w = h->weights;
    a.w1 = w; a.b1 = a.w1 + hid * od; a.w2 = a.b1 + hid; a.b2 = a.w2 + hid * hid; a.w3 = a.b2 + hid; a.b3 = a.w3 + hid;
    a.obs_dim = od;
}

}  // namespace

extern "C" {

const char *b2p_last_error(b2p_handle h) { return h ? h->error.c_str() : g_policy_create_error.c_str(); }

int b2p_create(int device, int obs_dim, int tanh_mode, b2p_handle *out) {
    if (!out) return pfail(nullptr, "b2p_create: null output pointer");
    *out = nullptr;
    if (obs_dim < 1 || obs_dim > B2P_MAX_OBS_DIM) return pfail(nullptr, "b2p_create: obs_dim must be 1..15");
    if (tanh_mode < 0 || tanh_mode > 2) return pfail(nullptr, "b2p_create: unknown tanh_mode");
    PolicyDeviceGuard guard(device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return pfail(nullptr, "b2p_create: no such CUDA device");
    if (prop.major != 10) return pfail(nullptr, "b2p_create: the policy kernel needs an sm_100 GPU (tcgen05); there is no fallback");
    if (!(set_policy_smem<0>() && set_policy_smem<1>() && set_policy_smem<2>()))
        return pfail(nullptr, "b2p_create: the policy kernel does not fit shared memory");
    b2p_policy *h = new (std::nothrow) b2p_policy();
    if (!h) return pfail(nullptr, "b2p_create: out of host memory");
    h->device = device; h->obs_dim = obs_dim; h->tanh_mode = tanh_mode; h->num_sms = prop.multiProcessorCount;
    h->have_weights = false;
    h->seed_counter = nullptr;
    const size_t count = (size_t)pol::HID * obs_dim + pol::HID + pol::HID * pol::HID + pol::HID + pol::HID + 1;
    if (cudaMalloc((void **)&h->weights, count * sizeof(float)) != cudaSuccess) {
        delete h;
        return pfail(nullptr, "b2p_create: cudaMalloc failed");
    }
    *out = h;
    return 0;
}

int b2p_set_seed_counter(b2p_handle h, const uint64_t *counter) {
    if (!h) return 1;
    h->seed_counter = reinterpret_cast<const unsigned long long *>(counter);
    return 0;
}

void b2p_destroy(b2p_handle h) {
    if (!h) return;
    PolicyDeviceGuard guard(h->device);
    cudaFree(h->weights);
    delete h;
}

int b2p_set_weights(b2p_handle h, const float *w1, const float *b1, const float *w2, const float *b2,
                    const float *w3, const float *b3, void *stream) {
    if (!h) return 1;
    if (!w1 || !b1 || !w2 || !b2 || !w3 || !b3) return pfail(h, "b2p_set_weights: null pointer");
    PolicyDeviceGuard guard(h->device);
    const cudaStream_t cs = (cudaStream_t)stream;
    const int od = h->obs_dim, hid = pol::HID;
    float *w = h->weights;
    const float *src[6] = {w1, b1, w2, b2, w3, b3};
    const size_t cnt[6] = {(size_t)hid * od, (size_t)hid, (size_t)hid * hid, (size_t)hid, (size_t)hid, 1};
    for (int i = 0; i < 6; ++i) {
        if (cudaMemcpyAsync(w, src[i], cnt[i] * sizeof(float), cudaMemcpyDeviceToDevice, cs) != cudaSuccess)
            return pfail(h, std::string("b2p_set_weights: ") + cudaGetErrorString(cudaGetLastError()));
        w += cnt[i];
    }
    h->have_weights = true;
    return 0;
}

int b2p_act(b2p_handle h, const float *obs, int64_t rows, float *actions_out, float noise_std,
            uint64_t seed, float low, float high, void *stream) {
    if (!h) return 1;
    if (!obs || !actions_out || rows < 0) return pfail(h, "b2p_act: bad argument");
    if (!h->have_weights) return pfail(h, "b2p_act: b2p_set_weights has not been called");
    if (rows == 0) return 0;
    PolicyDeviceGuard guard(h->device);
    pol::Args a;
    memset(&a, 0, sizeof(a));
    fill_weights(h, a);
    a.obs = obs; a.out = actions_out; a.rows = rows; a.tiles = (rows + pol::TILE - 1) / pol::TILE;
    a.seed_counter = h->seed_counter;
    a.seed = seed; a.noise_std = noise_std; a.low = low; a.high = high;
    a.bulk_ok = (reinterpret_cast<uintptr_t>(obs) & 15u) == 0 ? 1 : 0;
    Dev d;
    memset(&d, 0, sizeof(d));
    const long long want = (a.tiles + pol::GROUPS - 1) / pol::GROUPS;
    const int grid = (int)(want < h->num_sms ? want : h->num_sms);
    const cudaError_t err = launch_policy<0>(h->tanh_mode, grid, (cudaStream_t)stream, a, d);
    if (err != cudaSuccess) return pfail(h, std::string("b2p_act: ") + cudaGetErrorString(err));
    return 0;
}

int b2p_act_env(b2p_handle h, b2e_handle env, float *actions_out, float noise_std, uint64_t seed,
                float low, float high, void *stream) {
    if (!h) return 1;
    if (!env || !actions_out) return pfail(h, "b2p_act_env: null pointer");
    if (!h->have_weights) return pfail(h, "b2p_act_env: b2p_set_weights has not been called");
    int ring_ok = 0, device = -1;
    const Dev *dv = static_cast<const Dev *>(b2e_dev_view(env, &ring_ok, &device));
    if (!dv || !ring_ok) return pfail(h, "b2p_act_env: the env has no MultiOptLRs adjusted-history rings (large-problem pipeline only)");
    if (device != h->device) return pfail(h, "b2p_act_env: policy and env live on different devices");
    if (3 * dv->H != h->obs_dim) return pfail(h, "b2p_act_env: obs_dim of the policy is not 3 x max_history of the env");
    PolicyDeviceGuard guard(h->device);
    pol::Args a;
    memset(&a, 0, sizeof(a));
    fill_weights(h, a);
    a.out = actions_out;
    a.rows = (long long)dv->E * dv->P;
    const int tiles_per_env = (dv->P + pol::TILE - 1) / pol::TILE;
    a.tiles = (long long)dv->E * tiles_per_env;
    // few envs: split every env's tiles into segments so that all SMs have work (>= 4 rounds of tiles per group and segment)
    int segs = (3 * h->num_sms + dv->E - 1) / dv->E;
    const int max_segs = tiles_per_env / (4 * pol::GROUPS);
    segs = segs > max_segs ? max_segs : segs;
    a.segs = segs < 1 ? 1 : segs;
    a.units = dv->E * a.segs;
    a.seed_counter = h->seed_counter;
    a.seed = seed; a.noise_std = noise_std; a.low = low; a.high = high;
    const int grid = a.units < h->num_sms ? a.units : h->num_sms;
    const cudaError_t err = dv->H == 5 ? launch_policy<1>(h->tanh_mode, grid, (cudaStream_t)stream, a, *dv)
                                      : launch_policy<2>(h->tanh_mode, grid, (cudaStream_t)stream, a, *dv);
    if (err != cudaSuccess) return pfail(h, std::string("b2p_act_env: ") + cudaGetErrorString(err));
    return 0;
}

}  // extern "C"

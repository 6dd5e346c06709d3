// b200data.cu -- device data front-end of libb200env.so (C ABI in include/b200data.h), sm_100a.
//
// Restates, on the device, what custom_envs/data/load_data.py:47-112 does on the host once per
// data set: Pillow NEAREST down-sampling (utils/utils_image.py:6-24), min-max normalisation in
// float64 (utils/utils_math.py:77-87) and label ranking / one-hot (utils/utils_common.py:88-99).
// Everything here is streaming byte / word work: grids are sized in multiples of the SM count,
// loads and stores are coalesced along the row-major inner dimension, reductions are two-pass
// with a fixed order (no atomics on values), so results do not depend on the launch shape.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "b200data.h"

namespace {

thread_local char g_error[256] = "";

int fail(int code, const char* what, cudaError_t err = cudaSuccess) {
    if (err != cudaSuccess)
        snprintf(g_error, sizeof(g_error), "%s: %s", what, cudaGetErrorString(err));
    else
        snprintf(g_error, sizeof(g_error), "%s", what);
    return code;
}

int sm_count() {
    static int count = 0;
    if (!count) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || count <= 0)
            count = 148;
    }
    return count;
}

// ------------------------------------------------------------------ MT19937 shuffle (numpy legacy RandomState)
// One lane per generator; the 624-word state of the warp's 32 generators sits in shared memory word-interleaved
// ([624][32]: lane l touches bank l only).  mt19937_gen / mt19937_next / random_interval / _shuffle_raw of
// numpy/random/src/mt19937 and numpy/random/_common restated.
constexpr int kMtN = 624, kMtM = 397;

__device__ __forceinline__ void mt_twist(uint32_t* mt, int lane) {       // regenerate all 624 words
    auto at = [&](int i) -> uint32_t& { return mt[i * 32 + lane]; };
    int kk = 0;
    for (; kk < kMtN - kMtM; ++kk) {
        const uint32_t y = (at(kk) & 0x80000000u) | (at(kk + 1) & 0x7fffffffu);
        at(kk) = at(kk + kMtM) ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < kMtN - 1; ++kk) {
        const uint32_t y = (at(kk) & 0x80000000u) | (at(kk + 1) & 0x7fffffffu);
        at(kk) = at(kk + (kMtM - kMtN)) ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    const uint32_t y = (at(kMtN - 1) & 0x80000000u) | (at(0) & 0x7fffffffu);
    at(kMtN - 1) = at(kMtM - 1) ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

__global__ void __launch_bounds__(32) mt_shuffle_kernel(const uint32_t* __restrict__ gen, int mode, int64_t count, int n,
                                                         int32_t* __restrict__ out) {
    extern __shared__ uint32_t mt[];                                       // [624][32]
    const int lane = threadIdx.x;
    const int64_t g = (int64_t)blockIdx.x * 32 + lane;
    const bool live = g < count;
    int pos = kMtN;
    if (live) {
        if (mode == B2D_MT_SEEDS) {                                        // mt19937_seed
            uint32_t seed = gen[g];
            for (int i = 0; i < kMtN; ++i) {
                mt[i * 32 + lane] = seed;
                seed = 1812433253u * (seed ^ (seed >> 30)) + (uint32_t)i + 1u;
            }
        } else {
            const uint32_t* st = gen + g * (kMtN + 1);
            for (int i = 0; i < kMtN; ++i) mt[i * 32 + lane] = st[i];
            pos = (int)st[kMtN];
        }
        int32_t* x = out + g * n;
        for (int i = 0; i < n; ++i) x[i] = i;
        for (int i = n - 1; i >= 1; --i) {                                 // _shuffle_raw
            uint32_t mask = (uint32_t)i;                                   // random_interval(i): smallest 2^k - 1 >= i
            mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
            uint32_t j;
            do {
                if (pos >= kMtN) { mt_twist(mt, lane); pos = 0; }
                uint32_t y = mt[pos * 32 + lane];
                ++pos;
                y ^= y >> 11;                                              // tempering
                y ^= (y << 7) & 0x9d2c5680u;
                y ^= (y << 15) & 0xefc60000u;
                y ^= y >> 18;
                j = y & mask;
            } while (j > (uint32_t)i);
            const int32_t t = x[i];
            x[i] = x[j];
            x[j] = t;
        }
    }
}

constexpr int kThreads = 256;
constexpr int kMinmaxBlocksPerSm = 4;
constexpr int kMaxLabel = 65536;

template <typename T>
__device__ __forceinline__ double as_double(T v) { return static_cast<double>(v); }

// ---------------------------------------------------------------- nearest resize
// One thread per output pixel; consecutive threads write consecutive output bytes/words and read
// one source row segment per output row.  The tables sit in shared memory.
template <typename T>
__global__ void __launch_bounds__(kThreads)
resize_nearest_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t count, int src_h, int src_w,
                      const int32_t* __restrict__ ytab, const int32_t* __restrict__ xtab, int dst_h, int dst_w) {
    extern __shared__ int32_t src_of[];             // [dst_h*dst_w] offset inside the source image
    const int dst_px = dst_h * dst_w;
    for (int i = threadIdx.x; i < dst_px; i += blockDim.x)
        src_of[i] = ytab[i / dst_w] * src_w + xtab[i % dst_w];
    __syncthreads();
    const int64_t src_px = int64_t(src_h) * src_w;
    const int64_t total = count * dst_px;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t image = i / dst_px;
        const int px = int(i - image * dst_px);
        out[i] = in[image * src_px + src_of[px]];
    }
}

// ---------------------------------------------------------------- column min / max
// Pass 1: CTA b owns a contiguous band of rows.  Narrow rows (cols divides the CTA size): the
// band is read as one flat element stream, a thread's column never changes and its extrema stay
// in registers.  Otherwise thread t owns columns t, t+256, ... and walks the band's rows, so a
// warp reads 32 consecutive elements of a row.  Per-thread extrema meet per column in shared
// memory in thread order, per-CTA results in pass 2 in CTA order (no value atomics).
template <typename T>
__global__ void __launch_bounds__(kThreads)
minmax_partial_kernel(const T* __restrict__ data, int64_t rows, int cols, int64_t rows_per_block,
                      double* __restrict__ part_min, double* __restrict__ part_max) {
    extern __shared__ double smem[];                 // [2][cols]
    double* smin = smem;
    double* smax = smem + cols;
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    for (int c = threadIdx.x; c < cols; c += blockDim.x) { smin[c] = inf; smax[c] = -inf; }
    __syncthreads();
    const int64_t row0 = int64_t(blockIdx.x) * rows_per_block;
    int64_t row1 = row0 + rows_per_block;
    if (row1 > rows) row1 = rows;
    if (row0 < row1) {
        if (cols <= blockDim.x && blockDim.x % cols == 0) {
            // narrow rows, column fixed per thread: keep the extrema in registers
            const int c = threadIdx.x % cols;
            double lo = inf, hi = -inf;
            const T* base = data + row0 * cols;
            const int64_t count = (row1 - row0) * cols;
            for (int64_t i = threadIdx.x; i < count; i += blockDim.x) {
                const double v = as_double(base[i]);
                lo = v < lo ? v : lo;
                hi = v > hi ? v : hi;
            }
            // merge the blockDim/cols threads of a column in thread order
            for (int turn = 0; turn < blockDim.x / cols; ++turn) {
                if (threadIdx.x / cols == turn) {
                    if (lo < smin[c]) smin[c] = lo;
                    if (hi > smax[c]) smax[c] = hi;
                }
                __syncthreads();
            }
        } else {
            // general: thread t owns columns t, t+blockDim, ... and walks the band's rows;
            // consecutive threads read consecutive columns of one row
            for (int c = threadIdx.x; c < cols; c += blockDim.x) {
                double lo = inf, hi = -inf;
                for (int64_t r = row0; r < row1; ++r) {
                    const double v = as_double(data[r * cols + c]);
                    lo = v < lo ? v : lo;
                    hi = v > hi ? v : hi;
                }
                smin[c] = lo;
                smax[c] = hi;
            }
            __syncthreads();
        }
    }
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        part_min[int64_t(blockIdx.x) * cols + c] = smin[c];
        part_max[int64_t(blockIdx.x) * cols + c] = smax[c];
    }
}

// Pass 2: one thread per column folds the per-CTA partials in CTA order.
__global__ void minmax_final_kernel(const double* __restrict__ part_min, const double* __restrict__ part_max,
                                    int blocks, int cols, double* __restrict__ mins, double* __restrict__ maxes) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double lo = part_min[c], hi = part_max[c];
    for (int b = 1; b < blocks; ++b) {
        const double l = part_min[int64_t(b) * cols + c], h = part_max[int64_t(b) * cols + c];
        lo = l < lo ? l : lo;
        hi = h > hi ? h : hi;
    }
    mins[c] = lo;
    maxes[c] = hi;
}

// ---------------------------------------------------------------- normalise
// One thread per element of the dense [rows][cols] input; float64 arithmetic, IEEE division.
// numexpr casts like C: integer and double columns are evaluated in float64 throughout, float32
// operands stay float32 until they meet the double literal 1e-8 / the double denominator.
template <typename T>
struct Expr {
    static __device__ __forceinline__ double num(T x, double mn) { return static_cast<double>(x) - mn; }
    static __device__ __forceinline__ double den(double mn, double mx) { return mx - mn + 1e-8; }
};
template <>
struct Expr<float> {
    static __device__ __forceinline__ double num(float x, double mn) { return static_cast<double>(x - static_cast<float>(mn)); }
    static __device__ __forceinline__ double den(double mn, double mx) {
        return static_cast<double>(static_cast<float>(mx) - static_cast<float>(mn)) + 1e-8;
    }
};

template <typename T, typename Out>
__global__ void __launch_bounds__(kThreads)
normalize_kernel(const T* __restrict__ data, int64_t rows, int cols, const double* __restrict__ mins,
                 const double* __restrict__ maxes, Out* __restrict__ out, int64_t out_stride) {
    extern __shared__ double smem[];                 // [2][cols]: min, denominator
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        smem[c] = mins[c];
        smem[cols + c] = Expr<T>::den(mins[c], maxes[c]);
    }
    __syncthreads();
    const int64_t total = rows * cols;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t r = i / cols;
        const int c = int(i - r * cols);
        const double v = Expr<T>::num(data[i], smem[c]) / smem[cols + c];
        out[r * out_stride + c] = static_cast<Out>(v);
    }
}

// ---------------------------------------------------------------- label ranks / one-hot
// workspace: int32 present[kMaxLabel], int32 rank_of[kMaxLabel], int32 flags[2] (bad, unique)
__global__ void mark_labels_kernel(const int32_t* __restrict__ labels, int64_t count, int32_t* __restrict__ present,
                                   int32_t* __restrict__ flags) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
        const int32_t v = labels[i];
        if (v < 0 || v >= kMaxLabel) flags[0] = 1;
        else if (!present[v]) present[v] = 1;        // every writer stores the same value
    }
}

// One CTA of 1024 threads: exclusive scan of present[] (64 values per thread).
__global__ void __launch_bounds__(1024)
rank_scan_kernel(const int32_t* __restrict__ present, int32_t* __restrict__ rank_of, int32_t* __restrict__ flags) {
    __shared__ int32_t warp_sums[32];
    constexpr int kPer = kMaxLabel / 1024;
    const int base = threadIdx.x * kPer;
    int32_t local = 0;
    for (int k = 0; k < kPer; ++k) local += present[base + k];
    int32_t incl = local;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = 1; d < 32; d <<= 1) {
        const int32_t up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int32_t w = warp_sums[lane];
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t up = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += up;
        }
        warp_sums[lane] = w;                         // inclusive over warps
    }
    __syncthreads();
    int32_t run = incl - local + (warp ? warp_sums[warp - 1] : 0);
    for (int k = 0; k < kPer; ++k) {
        rank_of[base + k] = run;
        run += present[base + k];
    }
    if (threadIdx.x == 1023) flags[1] = run;
}

__global__ void apply_ranks_kernel(const int32_t* __restrict__ labels, int64_t count, const int32_t* __restrict__ rank_of,
                                   int32_t* __restrict__ ranks) {
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += int64_t(gridDim.x) * blockDim.x) {
        const int32_t v = labels[i];
        ranks[i] = (v >= 0 && v < kMaxLabel) ? rank_of[v] : -1;
    }
}

template <typename Out>
__global__ void __launch_bounds__(kThreads)
onehot_kernel(const int32_t* __restrict__ ranks, int64_t count, int num_labels, Out* __restrict__ out,
              int32_t* __restrict__ flags) {
    const int64_t total = count * num_labels;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t r = i / num_labels;
        const int j = int(i - r * num_labels);
        const int32_t k = ranks[r];
        if (j == 0 && (k < 0 || k >= num_labels)) flags[0] = 1;
        out[i] = static_cast<Out>(k == j ? 1 : 0);
    }
}

int grid_for(int64_t work_items, int per_sm) {
    const int64_t want = (work_items + kThreads - 1) / kThreads;
    const int64_t cap = int64_t(sm_count()) * per_sm;
    return int(want < 1 ? 1 : (want < cap ? want : cap));
}

size_t elem_size(int dtype) {
    switch (dtype) {
        case B2D_U8: return 1;
        case B2D_I32: case B2D_F32: return 4;
        case B2D_F64: return 8;
        default: return 0;
    }
}

int minmax_blocks() { return sm_count() * kMinmaxBlocksPerSm; }

}  // namespace

extern "C" {

const char* b2d_last_error(void) { return g_error; }

int b2d_resize_nearest(const void* images, int dtype, int64_t count, int src_h, int src_w, const int32_t* ytab,
                       const int32_t* xtab, int dst_h, int dst_w, void* out, void* stream) {
    if (count < 0 || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || !elem_size(dtype))
        return fail(B2D_EINVAL, "b2d_resize_nearest: bad shape or dtype");
    if (count == 0) return B2D_OK;
    if (!images || !out || !ytab || !xtab) return fail(B2D_EINVAL, "b2d_resize_nearest: null pointer");
    const size_t smem = size_t(dst_h) * dst_w * sizeof(int32_t);
    if (smem > 48 * 1024) return fail(B2D_EINVAL, "b2d_resize_nearest: output image above 12288 pixels");
    auto s = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(count * dst_h * dst_w, 8);
    switch (dtype) {
        case B2D_U8:
            resize_nearest_kernel<uint8_t><<<grid, kThreads, smem, s>>>(
                static_cast<const uint8_t*>(images), static_cast<uint8_t*>(out), count, src_h, src_w, ytab, xtab, dst_h, dst_w);
            break;
        case B2D_I32: case B2D_F32:
            resize_nearest_kernel<uint32_t><<<grid, kThreads, smem, s>>>(
                static_cast<const uint32_t*>(images), static_cast<uint32_t*>(out), count, src_h, src_w, ytab, xtab, dst_h, dst_w);
            break;
        default:
            resize_nearest_kernel<uint64_t><<<grid, kThreads, smem, s>>>(
                static_cast<const uint64_t*>(images), static_cast<uint64_t*>(out), count, src_h, src_w, ytab, xtab, dst_h, dst_w);
    }
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? B2D_OK : fail(B2D_ECUDA, "b2d_resize_nearest", err);
}

int b2d_shuffle_permutations(const uint32_t* gen, int mode, int64_t count, int n, int32_t* out, void* stream) {
    if (count < 0 || n < 0 || (mode != B2D_MT_SEEDS && mode != B2D_MT_STATES))
        return fail(B2D_EINVAL, "b2d_shuffle_permutations: bad count, length or mode");
    if (count == 0 || n == 0) return B2D_OK;
    if (!gen || !out) return fail(B2D_EINVAL, "b2d_shuffle_permutations: null pointer");
    const size_t smem = size_t(kMtN) * 32 * sizeof(uint32_t);
    {   // the opt-in is per device: set it on every call (microseconds against a kernel of milliseconds)
        const cudaError_t err = cudaFuncSetAttribute(mt_shuffle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return fail(B2D_ECUDA, "b2d_shuffle_permutations: shared memory", err);
    }
    const int64_t blocks = (count + 31) / 32;
    if (blocks > 0x7fffffff) return fail(B2D_EINVAL, "b2d_shuffle_permutations: too many generators");
    mt_shuffle_kernel<<<(unsigned)blocks, 32, smem, static_cast<cudaStream_t>(stream)>>>(gen, mode, count, n, out);
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? B2D_OK : fail(B2D_ECUDA, "b2d_shuffle_permutations", err);
}

size_t b2d_minmax_workspace(int cols) {
    return cols <= 0 ? 0 : size_t(2) * minmax_blocks() * size_t(cols) * sizeof(double);
}

int b2d_column_minmax(const void* data, int dtype, int64_t rows, int cols, double* mins, double* maxes,
                      void* workspace, void* stream) {
    if (rows <= 0 || cols <= 0 || !elem_size(dtype))
        return fail(B2D_EINVAL, "b2d_column_minmax: needs at least one row and one column (numpy raises on empty)");
    if (!data || !mins || !maxes || !workspace) return fail(B2D_EINVAL, "b2d_column_minmax: null pointer");
    const size_t smem = size_t(2) * cols * sizeof(double);
    if (smem > 48 * 1024) return fail(B2D_EINVAL, "b2d_column_minmax: more than 3072 columns");
    auto s = static_cast<cudaStream_t>(stream);
    int blocks = minmax_blocks();
    int64_t rows_per_block = (rows + blocks - 1) / blocks;
    if (rows_per_block < 1) rows_per_block = 1;
    blocks = int((rows + rows_per_block - 1) / rows_per_block);
    double* part_min = static_cast<double*>(workspace);
    double* part_max = part_min + size_t(minmax_blocks()) * cols;
#define B2D_MINMAX(T) minmax_partial_kernel<T><<<blocks, kThreads, smem, s>>>( \
        static_cast<const T*>(data), rows, cols, rows_per_block, part_min, part_max)
    switch (dtype) {
        case B2D_U8: B2D_MINMAX(uint8_t); break;
        case B2D_I32: B2D_MINMAX(int32_t); break;
        case B2D_F32: B2D_MINMAX(float); break;
        default: B2D_MINMAX(double);
    }
#undef B2D_MINMAX
    minmax_final_kernel<<<(cols + 127) / 128, 128, 0, s>>>(part_min, part_max, blocks, cols, mins, maxes);
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? B2D_OK : fail(B2D_ECUDA, "b2d_column_minmax", err);
}

int b2d_normalize(const void* data, int dtype, int64_t rows, int cols, const double* mins, const double* maxes,
                  void* out, int out_dtype, int64_t out_stride, void* stream) {
    if (rows < 0 || cols <= 0 || !elem_size(dtype) || out_stride < cols || (out_dtype != B2D_F32 && out_dtype != B2D_F64))
        return fail(B2D_EINVAL, "b2d_normalize: bad shape, stride or dtype");
    if (rows == 0) return B2D_OK;
    if (!data || !mins || !maxes || !out) return fail(B2D_EINVAL, "b2d_normalize: null pointer");
    const size_t smem = size_t(2) * cols * sizeof(double);
    if (smem > 48 * 1024) return fail(B2D_EINVAL, "b2d_normalize: more than 3072 columns");
    auto s = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(rows * cols, 8);
#define B2D_NORM(T, O) normalize_kernel<T, O><<<grid, kThreads, smem, s>>>( \
        static_cast<const T*>(data), rows, cols, mins, maxes, static_cast<O*>(out), out_stride)
#define B2D_NORM_IN(O) switch (dtype) { \
        case B2D_U8: B2D_NORM(uint8_t, O); break; \
        case B2D_I32: B2D_NORM(int32_t, O); break; \
        case B2D_F32: B2D_NORM(float, O); break; \
        default: B2D_NORM(double, O); }
    if (out_dtype == B2D_F32) { B2D_NORM_IN(float) } else { B2D_NORM_IN(double) }
#undef B2D_NORM_IN
#undef B2D_NORM
    const cudaError_t err = cudaGetLastError();
    return err == cudaSuccess ? B2D_OK : fail(B2D_ECUDA, "b2d_normalize", err);
}

size_t b2d_rank_workspace(void) { return (size_t(2) * kMaxLabel + 2) * sizeof(int32_t); }

int b2d_label_ranks(const int32_t* labels, int64_t count, int32_t* ranks, int32_t* num_unique_host,
                    void* workspace, void* stream) {
    if (count < 0 || !num_unique_host || !workspace || (count && (!labels || !ranks)))
        return fail(B2D_EINVAL, "b2d_label_ranks: bad arguments");
    auto s = static_cast<cudaStream_t>(stream);
    int32_t* present = static_cast<int32_t*>(workspace);
    int32_t* rank_of = present + kMaxLabel;
    int32_t* flags = rank_of + kMaxLabel;
    cudaError_t err = cudaMemsetAsync(workspace, 0, b2d_rank_workspace(), s);
    if (err != cudaSuccess) return fail(B2D_ECUDA, "b2d_label_ranks: memset", err);
    if (count) mark_labels_kernel<<<grid_for(count, 8), kThreads, 0, s>>>(labels, count, present, flags);
    rank_scan_kernel<<<1, 1024, 0, s>>>(present, rank_of, flags);
    if (count) apply_ranks_kernel<<<grid_for(count, 8), kThreads, 0, s>>>(labels, count, rank_of, ranks);
    int32_t host_flags[2] = {0, 0};
    err = cudaMemcpyAsync(host_flags, flags, sizeof(host_flags), cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    if (err != cudaSuccess) return fail(B2D_ECUDA, "b2d_label_ranks", err);
    if (host_flags[0]) return fail(B2D_ERANGE, "b2d_label_ranks: a label lies outside [0, 65536)");
    *num_unique_host = host_flags[1];
    return B2D_OK;
}

int b2d_onehot(const int32_t* ranks, int64_t count, int num_labels, void* out, int out_dtype, void* workspace,
               void* stream) {
    if (count < 0 || num_labels <= 0 || !workspace || (out_dtype != B2D_F32 && out_dtype != B2D_F64))
        return fail(B2D_EINVAL, "b2d_onehot: bad arguments");
    if (count == 0) return B2D_OK;
    if (!ranks || !out) return fail(B2D_EINVAL, "b2d_onehot: null pointer");
    auto s = static_cast<cudaStream_t>(stream);
    int32_t* flags = static_cast<int32_t*>(workspace) + 2 * kMaxLabel;
    cudaError_t err = cudaMemsetAsync(flags, 0, 2 * sizeof(int32_t), s);
    if (err != cudaSuccess) return fail(B2D_ECUDA, "b2d_onehot: memset", err);
    const int grid = grid_for(count * num_labels, 8);
    if (out_dtype == B2D_F32)
        onehot_kernel<float><<<grid, kThreads, 0, s>>>(ranks, count, num_labels, static_cast<float*>(out), flags);
    else
        onehot_kernel<double><<<grid, kThreads, 0, s>>>(ranks, count, num_labels, static_cast<double*>(out), flags);
    int32_t bad = 0;
    err = cudaMemcpyAsync(&bad, flags, sizeof(bad), cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    if (err != cudaSuccess) return fail(B2D_ECUDA, "b2d_onehot", err);
    if (bad) return fail(B2D_ERANGE, "b2d_onehot: a label rank is >= num_labels (numpy raises IndexError)");
    return B2D_OK;
}

}  // extern "C"

/* b200data.h -- C ABI of the device data front-end of libb200env.so.
 *
 * What the reference does on the host, once per data set, before any env exists
 * (custom_envs/data/load_data.py:47-112): down-sample the images with Pillow's NEAREST filter
 * (utils/utils_image.py:6-24), min-max normalise every feature column
 * (utils/utils_math.py:77-87), turn the labels into ranks among their unique values and a
 * one-hot matrix (utils/utils_common.py:88-99).  Here the raw bytes are uploaded once and these
 * steps run on the device; their outputs (float32 feature rows, int32 label ranks) are what
 * b2e_bind_dataset (b200env.h) takes, so the data set never visits the host in float form.
 *
 * Every pointer is a DEVICE pointer unless it says "host".  `stream` is a cudaStream_t passed as
 * void* (NULL = the default stream).  Functions return 0 or a B2D_E* code; b2d_last_error() gives
 * the text of this thread's last failure.  Calls are asynchronous on `stream` except where noted.
 */
#ifndef B200DATA_H
#define B200DATA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { B2D_OK = 0, B2D_EINVAL = 1, B2D_ECUDA = 2, B2D_ERANGE = 3 };
/* element types of the raw data */
enum { B2D_U8 = 0, B2D_I32 = 1, B2D_F32 = 2, B2D_F64 = 3 };

const char* b2d_last_error(void);

/* Nearest-neighbour down-sampling of `count` row-major single-channel images, as
 * PIL.Image.resize(size, resample=0) does it (utils_image.py:6-14 -> Pillow ImagingScaleAffine):
 * out[n][y][x] = in[n][ytab[y]][xtab[x]].  The two source-index tables are DEVICE arrays of
 * dst_h / dst_w int32 built by the caller with Pillow's own accumulation (the host mirror
 * does that: custom_envs_b200/utils/utils_image.py); entries must lie inside the source. */
int b2d_resize_nearest(const void* images, int dtype, int64_t count, int src_h, int src_w,
                       const int32_t* ytab, const int32_t* xtab, int dst_h, int dst_w,
                       void* out, void* stream);

/* Column minima and maxima of a row-major [rows][cols] matrix (np.min/np.max over axis 0,
 * utils_math.py:83-84), as doubles.  `workspace` holds b2d_minmax_workspace(cols) bytes. */
size_t b2d_minmax_workspace(int cols);
int b2d_column_minmax(const void* data, int dtype, int64_t rows, int cols,
                      double* mins, double* maxes, void* workspace, void* stream);

/* out[r][c] = (data[r][c] - mins[c]) / (maxes[c] - mins[c] + 1e-8), evaluated in float64 as
 * numexpr does (utils_math.py:85-87) and stored as float32 (out_dtype = B2D_F32, what the
 * problem's float32 feed makes of it) or float64.  `out_stride` = elements between output rows
 * (>= cols; columns cols..out_stride-1 are left untouched). */
int b2d_normalize(const void* data, int dtype, int64_t rows, int cols, const double* mins,
                  const double* maxes, void* out, int out_dtype, int64_t out_stride, void* stream);

/* Labels -> ranks among their sorted unique values (np.unique(..., return_inverse=True),
 * utils_common.py:94).  Labels are int32 in [0, 65536).  Writes ranks[count] and the number of
 * distinct labels to *num_unique_host (HOST int; the call synchronises `stream`).
 * `workspace` holds b2d_rank_workspace() bytes.  B2D_ERANGE if a label is out of range. */
size_t b2d_rank_workspace(void);
int b2d_label_ranks(const int32_t* labels, int64_t count, int32_t* ranks, int32_t* num_unique_host,
                    void* workspace, void* stream);

/* onehot[i][j] = (ranks[i] == j), [count][num_labels], float32 or float64
 * (utils_common.py:97-98).  B2D_ERANGE (after synchronising) if a rank >= num_labels, where
 * numpy raises IndexError. */
int b2d_onehot(const int32_t* ranks, int64_t count, int num_labels, void* out, int out_dtype,
               void* workspace, void* stream);

/* The permutation `np.random.RandomState(seed).shuffle(np.arange(n))` yields -- what every epoch-end / reset shuffle
 * of one env applies to its data set (utils/utils_math.py:10-22 hands each of them the same generator state;
 * utils/utils_common.py:12-23; dataset/inmemorydataset.py:24-28) -- for `count` generators at once, bit for bit:
 * MT19937 as numpy's legacy RandomState drives it (init_genrand seeding, tempered 32-bit draws, masked rejection
 * sampling of random_interval, Fisher-Yates from the last element down).
 *   mode B2D_MT_SEEDS : `gen` = uint32 [count] integer seeds (RandomState(seed), 0 <= seed < 2^32)
 *   mode B2D_MT_STATES: `gen` = uint32 [count][625], the 624-word key then the position, i.e. RandomState.get_state()[1:3]
 *                       (a generator seeded any other way, e.g. classic gym's hashed seeds)
 * out: int32 [count][n] device.  One thread per generator (the shuffle is a dependent chain of n - 1 swaps). */
enum { B2D_MT_SEEDS = 0, B2D_MT_STATES = 1 };
int b2d_shuffle_permutations(const uint32_t* gen, int mode, int64_t count, int n, int32_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif

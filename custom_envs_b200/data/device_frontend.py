"""Device data front-end: the reference's per-data-set preparation (data/load_data.py:47-112)
as CUDA kernels over device-resident tensors (C ABI: include/b200data.h).

    images  uint8 [N, h, w]  --resize_nearest-->  uint8 [N, h', w']        (utils_image.py:6-24)
    table   [N, D]           --normalize-------->  float32 / float64 [N, D] (utils_math.py:77-87)
    labels  [N]              --label_ranks------>  int32 [N], count         (utils_common.py:94)
    ranks   [N]              --onehot----------->  float [N, C]             (utils_common.py:97-98)

The outputs are what ``BatchedOptEnv`` binds (float32 features, int32 label ranks), so a data set
uploaded as raw bytes never returns to the host.  No CPU fallback: every function raises
``B200EnvError`` when libb200env.so or a CUDA device is missing."""
import ctypes

import numpy as np
import torch

from custom_envs_b200 import _lib

_DTYPES = {torch.uint8: _lib.DTYPE_U8, torch.int32: _lib.DTYPE_I32, torch.float32: _lib.DTYPE_F32,
           torch.float64: _lib.DTYPE_F64}


def pillow_nearest_table(src_size, dst_size):
    """Source index per destination index of Pillow's NEAREST resize over the whole image
    (ImagingScaleAffine: start at scale/2, accumulate scale in float64, truncate)."""
    scale = float(src_size) / float(dst_size)
    pos, table = scale * 0.5, np.empty(dst_size, np.int32)
    for i in range(dst_size):
        table[i] = int(pos)
        pos += scale
    return table


def _lib_and_stream(tensor):
    lib = _lib.load()
    if not tensor.is_cuda:
        raise _lib.B200EnvError('the data front-end runs on CUDA tensors only (no CPU fallback)')
    return lib, ctypes.c_void_p(torch.cuda.current_stream(tensor.device).cuda_stream)


def _check(lib, code):
    if code:
        raise _lib.B200EnvError(lib.b2d_last_error().decode())


def _ptr(tensor):
    return ctypes.c_void_p(tensor.data_ptr())


def _as_device(array, device, dtypes=_DTYPES):
    tensor = array if torch.is_tensor(array) else torch.from_numpy(np.ascontiguousarray(array))
    if tensor.dtype == torch.int64 or tensor.dtype == torch.int16 or tensor.dtype == torch.int8:
        tensor = tensor.to(torch.int32)
    if tensor.dtype == torch.bool:
        tensor = tensor.to(torch.uint8)
    if tensor.dtype not in dtypes:
        raise _lib.B200EnvError('unsupported element type %s' % tensor.dtype)
    return tensor.to(device).contiguous()


def resize_nearest(images, shape, device='cuda:0'):
    """``resize_array_many`` for a stack [N, h, w]; ``shape`` is Pillow's (width, height)."""
    images = _as_device(images, device)
    assert images.dim() == 3
    lib, stream = _lib_and_stream(images)
    count, src_h, src_w = images.shape
    dst_w, dst_h = int(shape[0]), int(shape[1])
    ytab = torch.from_numpy(pillow_nearest_table(src_h, dst_h)).to(images.device)
    xtab = torch.from_numpy(pillow_nearest_table(src_w, dst_w)).to(images.device)
    out = torch.empty((count, dst_h, dst_w), dtype=images.dtype, device=images.device)
    with torch.cuda.device(images.device):
        _check(lib, lib.b2d_resize_nearest(_ptr(images), _DTYPES[images.dtype], count, src_h, src_w,
                                           _ptr(ytab), _ptr(xtab), dst_h, dst_w, _ptr(out), stream))
    return out


def column_minmax(table, device='cuda:0'):
    table = _as_device(table, device)
    assert table.dim() == 2
    lib, stream = _lib_and_stream(table)
    rows, cols = table.shape
    mins = torch.empty(cols, dtype=torch.float64, device=table.device)
    maxes = torch.empty_like(mins)
    with torch.cuda.device(table.device):
        work = torch.empty(max(1, lib.b2d_minmax_workspace(cols)), dtype=torch.uint8, device=table.device)
        _check(lib, lib.b2d_column_minmax(_ptr(table), _DTYPES[table.dtype], rows, cols, _ptr(mins),
                                          _ptr(maxes), _ptr(work), stream))
    return mins, maxes


def normalize(table, out_dtype=torch.float32, device='cuda:0', out=None):
    """Min-max scaling of every column, evaluated in float64 like the reference; ``out`` may be
    a wider row-major buffer (e.g. rows padded to 16 bytes), only columns [0, D) are written."""
    table = _as_device(table, device)
    mins, maxes = column_minmax(table)
    lib, stream = _lib_and_stream(table)
    rows, cols = table.shape
    if out is None:
        out = torch.empty((rows, cols), dtype=out_dtype, device=table.device)
    assert out.dim() == 2 and out.shape[0] == rows and out.stride(1) == 1 and out.shape[1] >= cols
    with torch.cuda.device(table.device):
        _check(lib, lib.b2d_normalize(_ptr(table), _DTYPES[table.dtype], rows, cols, _ptr(mins), _ptr(maxes),
                                      _ptr(out), _DTYPES[out.dtype], out.stride(0), stream))
    return out


def label_ranks(labels, device='cuda:0'):
    """(ranks int32 [N], number of distinct labels); labels must be integers in [0, 65536)."""
    if not torch.is_tensor(labels):
        labels = np.asarray(labels)
        if labels.dtype.kind == 'f':
            if not np.array_equal(labels, np.floor(labels)):
                raise _lib.B200EnvError('label ranks on the device need integral labels')
            labels = labels.astype(np.int64)
    labels = _as_device(labels, device, {torch.int32: 0, torch.uint8: 0}).to(torch.int32).reshape(-1)
    lib, stream = _lib_and_stream(labels)
    ranks = torch.empty_like(labels)
    count = ctypes.c_int32(0)
    with torch.cuda.device(labels.device):
        work = torch.empty(lib.b2d_rank_workspace(), dtype=torch.uint8, device=labels.device)
        _check(lib, lib.b2d_label_ranks(_ptr(labels), labels.numel(), _ptr(ranks), ctypes.byref(count),
                                        _ptr(work), stream))
    return ranks, int(count.value)


def to_onehot(labels, num_of_labels=None, out_dtype=torch.float64, device='cuda:0'):
    """``utils_common.to_onehot`` on the device: (onehot [N, C], C)."""
    ranks, unique = label_ranks(labels, device)
    num = unique if num_of_labels is None else int(num_of_labels)
    lib, stream = _lib_and_stream(ranks)
    out = torch.empty((ranks.numel(), num), dtype=out_dtype, device=ranks.device)
    with torch.cuda.device(ranks.device):
        work = torch.empty(lib.b2d_rank_workspace(), dtype=torch.uint8, device=ranks.device)
        _check(lib, lib.b2d_onehot(_ptr(ranks), ranks.numel(), num, _ptr(out), _DTYPES[out_dtype],
                                   _ptr(work), stream))
    return out, num


def image_dataset(images, labels, shape=(7, 7), num_of_labels=None, device='cuda:0'):
    """The image branch of ``load_data`` (mnist / fashion / emnist, load_data.py:71-103) from raw
    bytes: images uint8 [N, h*w] or [N, h, w] -> float32 features [N, shape[0]*shape[1]] and
    int32 label ranks on the device, ready for ``BatchedOptEnv(features, targets)``.
    Returns (features, ranks, num_classes)."""
    images = _as_device(images, device)
    if images.dim() == 2:
        side = int(round(images.shape[1] ** 0.5))
        assert side * side == images.shape[1], 'flat images must be square'
        images = images.reshape(-1, side, side)
    small = resize_nearest(images, shape)
    features = normalize(small.reshape(small.shape[0], -1))
    ranks, unique = label_ranks(labels, device)
    num = unique if num_of_labels is None else int(num_of_labels)
    if unique > num:
        raise _lib.B200EnvError('more distinct labels (%d) than num_of_labels (%d)' % (unique, num))
    return features, ranks, num

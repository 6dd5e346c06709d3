"""CPU restatement of the reference's data front-end -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may import this
module; the product path (custom_envs_b200/data/device_frontend.py) runs CUDA kernels and
raises without them.

Pinned by ``tests/golden/data_frontend.npz``: outputs of the reference's own
``utils_image.resize_array_many`` (running the real Pillow 12.2), ``utils_common.to_onehot`` and
``utils_math.normalize`` recorded by ``tests/golden/gen_data_golden.py``.  ``numexpr`` is a
third-party dependency that is absent from the image and unpinned by the reference
(requirements list it without a version); its published evaluation rules for the one expression
on this path are restated in ``numexpr_promote``: uint8/int8/uint16/int16/bool operands are
promoted to int32, uint32 to int64, mixed int/double arithmetic is carried out in float64, and
float32 operands stay float32 until they meet the double literal (C-style casting).
"""
import numpy as np


def pillow_nearest_table(src_size, dst_size):
    """Source index of every destination index for Image.resize(..., resample=NEAREST) over
    the full image: Pillow's ImagingScaleAffine starts at ``scale * 0.5`` and ACCUMULATES
    ``scale`` in float64, truncating towards zero (reference utils/utils_image.py:6-14 calls
    ``Image.fromarray(array).resize(shape, 0)``)."""
    scale = float(src_size) / float(dst_size)
    pos = 0.0 + scale * 0.5
    table = np.empty(dst_size, np.int32)
    for i in range(dst_size):
        table[i] = -1 if pos < 0.0 else int(pos)
        pos += scale
    return table


def resize_nearest(images, shape):
    """``resize_array_many`` (utils_image.py:17-24) for a stack [N, h, w]; ``shape`` is Pillow's
    (width, height).  Returns [N, shape[1], shape[0]]."""
    images = np.asarray(images)
    xtab = pillow_nearest_table(images.shape[2], shape[0])
    ytab = pillow_nearest_table(images.shape[1], shape[1])
    return images[:, ytab[:, None], xtab[None, :]]


def numexpr_promote(array):
    array = np.asarray(array)
    if array.dtype.kind in 'bui' and array.dtype.itemsize < 4:
        return array.astype(np.int32)
    if array.dtype == np.uint32:
        return array.astype(np.int64)
    return array


def normalize(data):
    """utils_math.py:77-87: (data - mins) / (maxes - mins + 1e-8), float64 result."""
    data = np.asarray(data)
    mins = numexpr_promote(np.min(data, axis=0))
    maxes = numexpr_promote(np.max(data, axis=0))
    data = numexpr_promote(data)
    # numexpr casts like C: float32 operands stay float32 until they meet the double literal
    return (data - mins) / (maxes - mins + np.float64(1e-8))


def to_onehot(array, num_of_labels=None):
    """utils_common.py:88-99: ranks among the sorted unique values, one-hot float64."""
    unique, inverse = np.unique(np.asarray(array), return_inverse=True)
    inverse = inverse.ravel()
    if num_of_labels is None:
        num_of_labels = unique.size
    onehot = np.zeros((len(inverse), num_of_labels))
    onehot[np.arange(len(inverse)), inverse] = 1
    return onehot, num_of_labels


def image_features(images, shape=(7, 7)):
    """The image branch of load_data (data/load_data.py:71-77): resize, flatten, normalise."""
    small = resize_nearest(images, shape)
    return normalize(small.reshape(len(small), -1))

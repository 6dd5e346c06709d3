"""Timing of the data front-end kernels on one GPU (CUDA events, best of 5 after warm-up)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from custom_envs_b200.data import device_frontend as dev          # noqa: E402


def timed(fn, reps=5):
    fn()
    best = 1e9
    for _ in range(reps):
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        fn()
        stop.record()
        torch.cuda.synchronize()
        best = min(best, start.elapsed_time(stop))
    return best


rng = np.random.RandomState(0)
rows = 60000
images = torch.from_numpy(rng.randint(0, 256, size=(rows, 28, 28)).astype(np.uint8)).cuda()
labels = torch.from_numpy(rng.randint(0, 10, rows).astype(np.int32)).cuda()
wide = torch.rand(rows, 784, device='cuda')
for name, fn, nbytes in [
        ('resize 60000x28x28 u8 -> 7x7', lambda: dev.resize_nearest(images, (7, 7)), rows * (784 // 4 + 49)),
        ('normalize 60000x49 u8 -> f32 (min/max + scale)',
         lambda: dev.normalize(images.reshape(rows, 784)[:, :49].contiguous()), rows * 49 * (1 + 1 + 4)),
        ('column_minmax 60000x784 f32', lambda: dev.column_minmax(wide), rows * 784 * 4),
        ('normalize 60000x784 f32 -> f32 (min/max + scale)', lambda: dev.normalize(wide), rows * 784 * 12),
        ('to_onehot 60000 labels -> [60000,10] f64', lambda: dev.to_onehot(labels), rows * (4 + 4 + 4 + 80)),
        ('image_dataset (resize + normalize + ranks)', lambda: dev.image_dataset(images, labels), 0)]:
    ms = timed(fn)
    print('%-52s %8.3f ms  %s' % (name, ms, ('%.0f GB/s' % (nbytes / ms / 1e6)) if nbytes else ''))

"""In-memory data set with the reference's batching rules (dataset/inmemorydataset.py:11-38):
batch ``k`` is rows ``[k*B, (k+1)*B)`` of the currently permuted arrays, the last batch is ragged,
``on_epoch_end`` permutes both arrays with one shuffle of the GLOBAL numpy RNG.  On the device
path only ``features`` / ``targets`` / ``batch_size`` are read (the kernels keep each env's row
order themselves); the keras ``Sequence``-like protocol is kept for host users."""
import collections

import numpy as np

BatchType = collections.namedtuple('BatchType', 'features labels')


class DataSet:
    """len / getitem / iteration / on_epoch_end."""

    def on_epoch_end(self):
        """Hook run by the consumer when an epoch's iterator is exhausted."""

    def __iter__(self):
        for k in range(len(self)):
            yield self[k]


class InMemoryDataSet(DataSet):
    def __init__(self, features, targets, batch_size=None):
        self.features, self.targets = np.asarray(features), np.asarray(targets)
        if len(self.features) != len(self.targets):
            raise AssertionError('features and targets differ in length')
        self.batch_size = int(batch_size) if batch_size is not None else len(self.features)

    feature_shape = property(lambda self: self.features.shape[1:])
    target_shape = property(lambda self: self.targets.shape[1:])

    def __len__(self):
        return -(-len(self.features) // self.batch_size)            # ceil

    def __getitem__(self, k):
        rows = slice(k * self.batch_size, (k + 1) * self.batch_size)
        return BatchType(self.features[rows], self.targets[rows])

    def on_epoch_end(self):
        order = np.arange(len(self.features))
        np.random.shuffle(order)                                    # global RNG, as the reference does
        self.features = self.features[order]
        self.targets = self.targets[order]


class DeviceDataSet(DataSet):
    """A data set prepared by the device front-end (custom_envs_b200/data/device_frontend.py):
    ``features`` float32 [N, D] and ``targets`` int32 label ranks [N] are CUDA tensors that the
    env kernels bind directly.  Batches are device slices; ``on_epoch_end`` is a no-op here
    because the kernels keep every env's row order themselves."""
    on_device = True

    def __init__(self, features, targets, batch_size=None, num_classes=None):
        if len(features) != len(targets):
            raise AssertionError('features and targets differ in length')
        self.features, self.targets = features, targets
        self.batch_size = int(batch_size) if batch_size is not None else len(features)
        self.num_classes = int(num_classes if num_classes is not None else int(targets.max().item()) + 1)

    feature_shape = property(lambda self: tuple(self.features.shape[1:]))
    target_shape = property(lambda self: (self.num_classes,))

    def __len__(self):
        return -(-len(self.features) // self.batch_size)

    def __getitem__(self, k):
        rows = slice(k * self.batch_size, (k + 1) * self.batch_size)
        return BatchType(self.features[rows], self.targets[rows])

"""In-memory data set with the reference's batching rules (dataset/inmemorydataset.py:11-38):
contiguous slices of ``batch_size`` rows, ragged last batch, whole-array reshuffle at epoch
end.  On the device path only ``features`` / ``targets`` / ``batch_size`` are read (the
kernel keeps each env's row order itself); the Sequence protocol is kept for host users."""
import math
from collections import namedtuple

import numpy as np

BatchType = namedtuple('BatchType', ['features', 'labels'])


class DataSet:
    """keras.utils.Sequence-like protocol: len / getitem / iteration / on_epoch_end."""

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def on_epoch_end(self):
        pass


class InMemoryDataSet(DataSet):
    def __init__(self, features, targets, batch_size=None):
        assert len(features) == len(targets)
        self.features = np.asarray(features)
        self.targets = np.asarray(targets)
        self.batch_size = len(self.features) if batch_size is None else int(batch_size)

    def on_epoch_end(self):
        order = np.arange(len(self.features))
        np.random.shuffle(order)                 # global RNG, as the reference does
        self.features, self.targets = self.features[order], self.targets[order]

    def __len__(self):
        return math.ceil(len(self.features) / self.batch_size)

    def __getitem__(self, idx):
        begin = idx * self.batch_size
        return BatchType(self.features[begin:begin + self.batch_size],
                         self.targets[begin:begin + self.batch_size])

    @property
    def feature_shape(self):
        return self.features.shape[1:]

    @property
    def target_shape(self):
        return self.targets.shape[1:]

"""Host containers kept for the wrappers and for API compatibility (reference
utils/utils_common.py).  On the device path the history lives in HBM rings; this numpy
``History`` serves ``HistoryWrapper`` and user code."""
from collections.abc import Mapping

import numpy as np
import numpy.random as npr


def shuffle(*args, np_random=npr):
    """Apply one random permutation to every argument (row alignment kept)."""
    order = np.arange(len(args[0]))
    np_random.shuffle(order)
    return [arg[order] for arg in args]


def to_onehot(array, num_of_labels=None):
    values, inverse = np.unique(array, return_inverse=True)
    if num_of_labels is None:
        num_of_labels = values.size
    onehot = np.zeros((len(inverse), num_of_labels))
    onehot[np.arange(len(inverse)), inverse.ravel()] = 1
    return onehot, num_of_labels


def flatten_arrays(arrays, dtype=np.float64):
    """Concatenate the ravelled arrays into one vector (float64 by default)."""
    if not arrays:
        return np.zeros(0, dtype)
    return np.concatenate([np.ravel(a) for a in arrays]).astype(dtype)


def from_flat(array, shapes):
    out, start = [], 0
    for shape in shapes:
        size = int(np.prod(shape))
        out.append(np.reshape(array[start:start + size], shape))
        start += size
    return out


def enzip(*iterables):
    for i, items in enumerate(zip(*iterables)):
        yield (i,) + items


class History(Mapping):
    """Fixed-depth history of named arrays; ``history[key]`` is ``[depth, *shape]`` with the
    NEWEST entry first; scalars (shape ``()``) are stored with shape ``(1,)``."""

    def __init__(self, max_history, **named_shapes):
        self.max_history = int(max_history)
        self.shapes = {name: tuple(shape) if shape else (1,) for name, shape in named_shapes.items()}
        self._data = {}
        self.iteration = 0
        self.reset()

    def __repr__(self):
        return '<History<max_history={}, shapes={!r}>>'.format(self.max_history, self.shapes)

    def __getitem__(self, key):
        # entries keep the dtype they were appended with (TensorFlow hands the env float32 losses):
        # the stacked array is float32 once every entry is, float64 while initial zeros remain
        return np.asarray(self._data[key])

    def __iter__(self):
        return iter(self.shapes)

    def __len__(self):
        return len(self.shapes)

    def reset_with_value(self, value):
        self._data = {name: [np.full(shape, value, np.float64)] * self.max_history
                      for name, shape in self.shapes.items()}
        self.iteration = 0

    def reset(self, **named_items):
        if named_items:
            assert self.keys() == named_items.keys()
            self._data = {name: [np.reshape(item, self.shapes[name])] * self.max_history
                          for name, item in named_items.items()}
            self.iteration = 0
        else:
            self.reset_with_value(0.0)

    def append(self, **named_items):
        assert self.keys() == named_items.keys()
        for name, item in named_items.items():
            self._data[name] = [np.reshape(item, self.shapes[name])] + self._data[name][:-1]   # newest first
        self.iteration = (self.iteration + 1) % self.max_history

    def build_multistate(self):
        """One tuple per agent: for each key (insertion order) its ``depth`` newest-first
        values; single-element keys are broadcast to every agent."""
        blocks = [self[name].reshape(self.max_history, -1) for name in self.shapes]
        width = max(block.shape[1] for block in blocks)
        rows = np.concatenate([np.broadcast_to(block, (self.max_history, width)) if block.shape[1] == 1
                               else block for block in blocks], axis=0)
        return [tuple(col) for col in rows.T.tolist()]

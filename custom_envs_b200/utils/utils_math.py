"""Host math helpers the boundary needs (reference utils/utils_math.py:10-22,77-87)."""
from contextlib import contextmanager

import numpy as np
import numpy.random as npr


@contextmanager
def use_random_state(random_state):
    """Run the body with the global numpy RNG loaded from a COPY of ``random_state``'s
    state; the global state is restored afterwards and ``random_state`` itself never
    advances (which is why each env reuses one permutation, see
    ``custom_envs_b200.batched_env.env_permutations``)."""
    saved = npr.get_state()
    try:
        npr.set_state(random_state.get_state())
        yield random_state
    finally:
        npr.set_state(saved)


def normalize(data):
    """Per-column min-max scaling to [0, 1]: (x - min) / (max - min + 1e-8)."""
    data = np.asarray(data, np.float64)
    mins, maxes = np.min(data, axis=0), np.max(data, axis=0)
    return (data - mins) / (maxes - mins + 1e-8)


def softmax(logits):
    logits = np.asarray(logits, np.float64)
    shifted = np.exp(logits - np.max(logits, axis=1)[:, None])
    return shifted / np.sum(shifted, axis=1)[:, None]


def cross_entropy(prob, ground_truth):
    return float(np.mean(np.sum(-np.log(np.asarray(prob) + 1e-16) * ground_truth, axis=1)))


def mse(prediction, ground_truth):
    return float(np.mean(np.sum((np.asarray(prediction) - ground_truth) ** 2, axis=1) / 2))

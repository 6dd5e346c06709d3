"""``load_data(name, batch_size)`` with the reference's normalisation and one-hot rules
(data/load_data.py:47-112).  The reference's data blobs are git-LFS pointers and there is
no network, so only ``iris`` (scikit-learn's bundled copy) and ``random_gaussians`` are
real; the MNIST-family names produce SYNTHETIC data of the reference's shape (49 features
after its 28x28 -> 7x7 down-sampling, 10 classes) and say so with a warning."""
import warnings

import numpy as np

from custom_envs_b200.dataset import InMemoryDataSet
from custom_envs_b200.utils.utils_common import to_onehot
from custom_envs_b200.utils.utils_math import normalize

_SYNTHETIC_ROWS = {'mnist': 60000, 'mnist-test': 10000, 'fashion': 60000,
                   'emnist-digits': 240000, 'cifar-10': 50000}


def synthetic_classification(num_rows, num_features, num_classes, seed=0):
    rng = np.random.RandomState(seed)
    features = normalize(rng.uniform(size=(num_rows, num_features)))
    labels = rng.randint(0, num_classes, size=num_rows)
    return features, labels


def _load_on_device(name, batch_size, num_of_labels, device):
    """Raw bytes / raw table -> device front-end -> ``DeviceDataSet`` (no float data on the host)."""
    from custom_envs_b200.data import device_frontend as front
    from custom_envs_b200.dataset import DeviceDataSet
    if name == 'iris':
        from sklearn import datasets
        iris = datasets.load_iris()
        features = front.normalize(iris.data, device=device)
        ranks, unique = front.label_ranks(iris.target, device=device)
    elif name in _SYNTHETIC_ROWS:
        warnings.warn('data set %r is not available offline; using synthetic 28x28 uint8 images '
                      'of its shape, down-sampled and normalised on the device' % name)
        rng = np.random.RandomState(0)
        images = rng.randint(0, 256, size=(_SYNTHETIC_ROWS[name], 28, 28)).astype(np.uint8)
        features, ranks, unique = front.image_dataset(images, rng.randint(0, 10, size=len(images)),
                                                      (7, 7), device=device)
    else:
        raise RuntimeError('No such data set named: {}'.format(name))
    num = unique if num_of_labels is None else int(num_of_labels)
    if unique > num:
        raise IndexError('more distinct labels (%d) than num_of_labels (%d)' % (unique, num))
    return DeviceDataSet(features, ranks, batch_size, num)


def load_data(name='iris', batch_size=32, num_of_labels=None, device=None):
    """``device=None`` prepares the arrays on the host exactly as the reference does;
    ``device='cuda:0'`` uploads the raw data and prepares it with the CUDA front-end."""
    if device is not None:
        return _load_on_device(name, batch_size, num_of_labels, device)
    if name == 'iris':
        from sklearn import datasets
        iris = datasets.load_iris()
        features = normalize(iris.data)
        labels, _ = to_onehot(iris.target, num_of_labels)
    elif name in _SYNTHETIC_ROWS:
        warnings.warn('data set %r is not available offline; using synthetic data of its '
                      'shape (49 features, 10 classes)' % name)
        features, raw = synthetic_classification(_SYNTHETIC_ROWS[name], 49, 10)
        labels, _ = to_onehot(raw, num_of_labels or 10)
    elif name == 'skin':
        warnings.warn("data set 'skin' is not available offline; using synthetic data")
        features3, raw = synthetic_classification(245057, 3, 2)
        features = np.zeros((len(features3), 4))
        features[:, :3] = features3
        labels, _ = to_onehot(raw, 3)
    elif name == 'random_gaussians':
        from sklearn import datasets
        features, raw = datasets.make_classification()
        labels, _ = to_onehot(raw, 2)
    else:
        raise RuntimeError('No such data set named: {}'.format(name))
    return InMemoryDataSet(features, labels, batch_size)

"""Device-backed problems: the BaseProblem API (problems/optimize_nn.py:122-159,
problems/optimize_function.py:97-137 in the reference) served by libb200env.so.

A problem is bound to one slot of a ``BatchedOptEnv``.  Used on its own it lazily creates
a private single-env backend; inside an env / ``OptVecEnv`` it is re-bound to the env's
slot of the shared backend, so ``env.model.get()`` reads the very state the kernel steps.
"""
import numpy as np

from custom_envs_b200.batched_env import BatchedOptEnv, ProblemSpec
from custom_envs_b200.problems.base_problem import BaseProblem


class _DeviceProblem(BaseProblem):
    spec = None
    data_set = None

    def __init__(self):
        self._backend = None
        self._slot = 0
        self._owns_backend = False
        self.device = 'cuda:0'

    # -- binding ---------------------------------------------------------------
    def device_arrays(self):
        """(features float32 [N,D], targets) as the kernel wants them, or (None, None)."""
        return None, None

    def data_key(self):
        return None

    @property
    def batch_size(self):
        return None

    def _bind(self, backend, slot, owns=False):
        if self._owns_backend and self._backend is not None and self._backend is not backend:
            self._backend.close()
        self._backend, self._slot, self._owns_backend = backend, slot, owns

    def _ensure(self):
        if self._backend is None:
            feats, targs = self.device_arrays()
            backend = BatchedOptEnv(self.spec, feats, targs, 1, batch_size=self.batch_size,
                                    row_order='natural', auto_reset=False, device=self.device)
            self._bind(backend, 0, owns=True)
            backend.reset()
        return self._backend

    def _mask(self):
        mask = np.zeros(self._ensure().num_envs, np.uint8)
        mask[self._slot] = 1
        return mask

    # -- BaseProblem -----------------------------------------------------------
    @property
    def size(self):
        return self.spec.size

    def reset(self):
        """Re-initialise the parameters and restart the data stream."""
        self._ensure().reset(env_mask=self._mask())

    def next(self):
        if self.spec.kind != 'func':
            self._ensure().next_batch(self._mask())

    def get(self):
        backend = self._ensure()
        grad, loss = backend.evaluate()
        params = backend.get_state('params')
        return (grad[self._slot].double().cpu().numpy(), np.float32(loss[self._slot].item()),
                params[self._slot].double().cpu().numpy())

    def get_gradient(self):
        return self.get()[0]

    def get_loss(self):
        return self.get()[1]

    def get_parameters(self):
        return self._ensure().get_state('params')[self._slot].double().cpu().numpy()

    def set_parameters(self, parameters):
        backend = self._ensure()
        params = backend.get_state('params')
        params[self._slot] = params.new_tensor(np.asarray(parameters, np.float64).astype(np.float32))
        backend.set_state('params', params)


class OptimizeNN(_DeviceProblem):
    """Dense stack + softmax cross-entropy on a data set (reference
    problems/optimize_nn.py:22-64).  ``layers`` are the hidden widths (relu); the reference
    builds them with ``model_fn`` / ``create_neural_net(layers=(256, 256))``.  Zero or one
    hidden layer runs the fused / streamed-operand kernels, deeper stacks the generic
    tiled-GEMM pipeline of libb200env.so."""

    def __init__(self, model_fn=None, data_set=None, layers=None):
        super().__init__()
        if data_set is None:
            from custom_envs_b200.data import load_data
            data_set = load_data()
        self.data_set = data_set
        if layers is None:
            layers = getattr(model_fn, 'layers', None)
        if layers is None:
            if model_fn is not None:
                raise NotImplementedError(
                    'OptimizeNN: arbitrary keras model_fn callables cannot be lowered to the '
                    'fused kernel; pass layers=(hidden_units,) or a model_fn with a .layers tuple')
            layers = (256, 256)                 # utils/utils_tf.py:74
        layers = tuple(int(h) for h in layers)
        if getattr(data_set, 'on_device', False):
            # prepared by the device front-end: float32 rows and int32 label ranks already in HBM
            self.spec = ProblemSpec('softmax', data_set.features.shape[1], layers, data_set.num_classes)
            self._features, self._labels = data_set.features, data_set.targets
            return
        features = np.asarray(data_set.features)
        targets = np.asarray(data_set.targets)
        num_outputs = targets.shape[1] if targets.ndim == 2 else int(targets.max()) + 1
        self.spec = ProblemSpec('softmax', features.shape[1], layers, num_outputs)
        self._labels = (targets.argmax(axis=1) if targets.ndim == 2 else targets).astype(np.int32)
        self._features = np.ascontiguousarray(features, np.float32)

    @classmethod
    def create(cls, model_fn=None, data_set=None, layers=None):
        return cls(model_fn, data_set, layers)

    def device_arrays(self):
        return self._features, self._labels

    def data_key(self):
        """Content digest of (features, labels): envs may share one data-set replica in HBM only if
        their arrays are EQUAL, not merely of equal shape and sum.  Computed once per data-set
        object and cached on it, so thousands of envs over one data set hash it once."""
        cached = getattr(self.data_set, '_b2e_digest', None)
        if cached is None:
            import hashlib
            digest = hashlib.blake2b(digest_size=16)
            for array in (self._features, self._labels):
                host = array.detach().cpu().numpy() if hasattr(array, 'detach') else np.asarray(array)
                digest.update(str((host.shape, host.dtype.str)).encode())
                digest.update(np.ascontiguousarray(host).tobytes())
            cached = digest.hexdigest()
            try:
                self.data_set._b2e_digest = cached
            except AttributeError:          # data-set type without a __dict__: recompute next time
                pass
        return cached

    @property
    def batch_size(self):
        return self.data_set.batch_size


class OptimizeFunction(_DeviceProblem):
    """Rosenbrock's function from (-1.9, 2.0) (reference problems/optimize_function.py:35-37)."""

    def __init__(self, function=None, initial_points=None, ndims=2):
        super().__init__()
        if function is not None or initial_points is not None or ndims != 2:
            raise NotImplementedError('OptimizeFunction: only the default Rosenbrock problem is built')
        self.spec = ProblemSpec('func', 0, (), 0)

    @classmethod
    def create(cls, function=None, initial_points=None, ndims=2):
        return cls(function, initial_points, ndims)

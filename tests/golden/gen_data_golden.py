"""Golden vectors of the data front-end, recorded from the REFERENCE'S OWN functions.

Run in the build container only (needs /root/reference and Pillow):

    python tests/golden/gen_data_golden.py

Executes, unmodified and imported by file path: ``custom_envs/utils/utils_image.py``
(``resize_array_many`` over the real Pillow), ``custom_envs/utils/utils_common.py``
(``to_onehot``) and ``custom_envs/utils/utils_math.py`` (``normalize``).  ``numexpr`` is absent
from the image: a stand-in ``evaluate`` applies numexpr's documented operand promotion
(small ints -> int32) and evaluates the expression with numpy in float64.
Writes tests/golden/data_frontend.npz; nothing under tests/ reads /root/reference at run time.
"""
import importlib.util
import os
import re
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = '/root/reference/custom_envs/utils'


def _numexpr_standin():
    module = types.ModuleType('numexpr')

    def promote(value):
        value = np.asarray(value)
        if value.dtype.kind in 'bui' and value.dtype.itemsize < 4:
            return value.astype(np.int32)
        return value

    def evaluate(expression, local_dict=None, **_):
        scope = {name: promote(value) for name, value in (local_dict or {}).items()}
        scope.update(exp=np.exp, sum=np.sum, _f64=np.float64)
        # numexpr casts like C: a double literal makes float32 operands double ("a*b returns a
        # float64 in Numexpr, but a float32 in NumPy", numexpr user guide, casting rules)
        expression = re.sub(r'(?<![\w.])(\d+\.?\d*e-?\d+|\d+\.\d*)', r'_f64(\1)', expression)
        return eval(expression, {'__builtins__': {}}, scope)       # noqa: S307 (fixed reference strings)

    module.evaluate = evaluate
    return module


def _load(name):
    spec = importlib.util.spec_from_file_location('ref_' + name, os.path.join(REFERENCE, name + '.py'))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


def main():
    sys.modules['numexpr'] = _numexpr_standin()
    image, common, math_ = _load('utils_image'), _load('utils_common'), _load('utils_math')
    rng = np.random.RandomState(7)
    out = {}
    # (height, width) of the source stack -> Pillow (width, height) of the result
    cases = [((28, 28), (7, 7)), ((28, 28), (10, 10)), ((32, 32), (7, 7)), ((5, 9), (4, 3)),
             ((28, 28), (28, 28)), ((13, 17), (5, 11))]
    for k, (src, dst) in enumerate(cases):
        stack = rng.randint(0, 256, size=(6,) + src).astype(np.uint8)
        small = np.stack(image.resize_array_many(list(stack), dst))
        out['resize%d_in' % k], out['resize%d_out' % k] = stack, small
        out['resize%d_shape' % k] = np.array(dst)
    # the mnist branch of load_data (data/load_data.py:71-77) on MNIST-shaped bytes
    digits = rng.randint(0, 256, size=(64, 784)).astype(np.uint8)
    digits[:, 58] = 9                                  # a constant column: 0 / 1e-8 = 0
    features = [f.reshape((28, 28)) for f in digits]
    features = np.reshape(image.resize_array_many(features, (7, 7)), (len(digits), -1))
    out['mnist_in'], out['mnist_small'] = digits, features
    out['mnist_features'] = math_.normalize(features)
    # float data (iris branch) and int32 data
    table = rng.normal(size=(150, 4)) * [1.0, 10.0, 0.1, 100.0] + [5.0, -3.0, 0.0, 50.0]
    out['float_in'], out['float_out'] = table, math_.normalize(table)
    table32 = table.astype(np.float32)
    out['float32_in'], out['float32_out'] = table32, math_.normalize(table32)
    ints = rng.randint(-1000, 1000, size=(97, 13)).astype(np.int32)
    out['int_in'], out['int_out'] = ints, math_.normalize(ints)
    # labels
    for k, (labels, count) in enumerate([(rng.randint(0, 10, 200), None), (rng.randint(0, 10, 200), 12),
                                          (np.array([5, 3, 3, 9, 5, 200]), None),
                                          (np.array([0.0, 1.0, 2.0, 1.0, 0.0]), 3)]):
        onehot, num = common.to_onehot(labels, count)
        out['labels%d_in' % k], out['labels%d_onehot' % k] = np.asarray(labels), onehot
        out['labels%d_num' % k] = np.array(num)
    np.savez_compressed(os.path.join(HERE, 'data_frontend.npz'), **out)
    print('wrote', len(out), 'arrays')


if __name__ == '__main__':
    main()

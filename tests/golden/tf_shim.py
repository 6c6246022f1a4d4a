"""Minimal numpy-backed stand-in for the slice of TensorFlow the reference's hot path touches.

TEST INFRASTRUCTURE ONLY.  TensorFlow is not installed in this image, so the reference's
own Python (`/root/reference/h2gcn/models/{__init__,_layers,H2GCN}.py`, `datasets/_dataset.py`) cannot run
as shipped.  `make_golden.py` installs this module as `sys.modules["tensorflow"]` and then imports and
executes the reference files UNMODIFIED; every op the reference calls on `tf.*` lands here.  The shim only
restates the *library* ops (third-party TensorFlow); the reference's own control flow — layer construction,
tag bookkeeping, hop stacking, concat order, adjacency powers, normalisation — is executed from the reference
sources themselves, which is what pins the golden vectors.

Restated TF semantics (from memory of TF 2.2; unverifiable offline, stated as assumptions):
  * tf.sparse.reorder            -> canonical row-major order (lexsort by row, then column)
  * tf.sparse.sparse_dense_matmul -> out zero-initialised; loop over nnz in stored order;
                                     out[row,:] += val * b[col,:] in fp32 (mul then add)
  * keras Dense / ReLU / Flatten / Dropout(training=False) / glorot_uniform default initialiser
"""
import re
import sys
import types

import numpy as np

_rng = np.random.default_rng(0)
_name_counts = {}


def seed(s):
    global _rng
    _rng = np.random.default_rng(s)
    _name_counts.clear()


class EagerTensor(np.ndarray):
    def numpy(self):
        return np.asarray(self)


def _wrap(a):
    return np.asarray(a).view(EagerTensor)


class SparseTensor:
    def __init__(self, indices, values, dense_shape):
        self.indices = _wrap(np.asarray(indices, dtype=np.int64).reshape(-1, 2))
        self.values = _wrap(np.asarray(values))
        self.dense_shape = _wrap(np.asarray(dense_shape, dtype=np.int64))

    @property
    def shape(self):
        return tuple(int(x) for x in self.dense_shape)

    @property
    def dtype(self):
        return self.values.dtype

    def __truediv__(self, s):
        return SparseTensor(self.indices, self.values / s, self.dense_shape)


def _reorder(sp):
    idx = np.asarray(sp.indices)
    order = np.lexsort((idx[:, 1], idx[:, 0]))
    return SparseTensor(idx[order], np.asarray(sp.values)[order], sp.dense_shape)


def _sparse_dense_matmul(sp_a, b):
    idx = np.asarray(sp_a.indices)
    vals = np.asarray(sp_a.values)
    b = np.asarray(b)
    out = np.zeros((sp_a.shape[0], b.shape[1]), dtype=np.result_type(vals.dtype, b.dtype))
    # ufunc.at is unbuffered and applies updates in index order => sequential accumulation in stored order
    step = max(1, (1 << 24) // max(1, b.shape[1]))
    for s in range(0, idx.shape[0], step):
        e = min(idx.shape[0], s + step)
        np.add.at(out, idx[s:e, 0], vals[s:e, None] * b[idx[s:e, 1]])
    return _wrap(out)


def _to_dense(sp):
    out = np.zeros(sp.shape, dtype=sp.values.dtype)
    idx = np.asarray(sp.indices)
    out[idx[:, 0], idx[:, 1]] = np.asarray(sp.values)
    return _wrap(out)


def _retain(sp, mask):
    mask = np.asarray(mask, dtype=bool)
    return SparseTensor(np.asarray(sp.indices)[mask], np.asarray(sp.values)[mask], sp.dense_shape)


def _snake(name):
    s = re.sub(r"(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub(r"([a-z0-9])([A-Z])", r"\1_\2", s).lower()


class Layer:
    def __init__(self, *a, **k):
        base = _snake(type(self).__name__)
        n = _name_counts.get(base, 0)
        _name_counts[base] = n + 1
        self.name = base if n == 0 else f"{base}_{n}"
        self.built = False
        self.weights = []

    def build(self, input_shape):
        self.built = True

    def add_weight(self, name=None, shape=None, regularizer=None, initializer=None, **k):
        shape = tuple(int(s) for s in shape)
        if initializer is None:  # keras default: glorot_uniform
            fan_in, fan_out = (shape[0], shape[-1]) if len(shape) > 1 else (shape[0], shape[0])
            limit = np.sqrt(6.0 / (fan_in + fan_out))
            w = _rng.uniform(-limit, limit, size=shape).astype(np.float32)
        else:
            w = np.zeros(shape, dtype=np.float32)
        w = _wrap(w)
        self.weights.append((name, w))
        return w

    def __call__(self, *args, **kwargs):
        if not self.built:
            self.build(getattr(args[0], "shape", None))
            self.built = True
        return self.call(*args, **kwargs)


class Model(Layer):
    def summary(self):
        pass


class ReLU(Layer):
    def call(self, x):
        return _wrap(np.maximum(np.asarray(x), 0))


class Flatten(Layer):
    def call(self, x):
        x = np.asarray(x)
        return _wrap(x.reshape(x.shape[0], -1))


class Dropout(Layer):
    def __init__(self, rate):
        super().__init__()
        self.rate = rate

    def call(self, x, training=False):
        return x  # inference only


class Dense(Layer):
    def __init__(self, units, use_bias=True, kernel_regularizer=None, **k):
        super().__init__()
        self.units = units
        self.use_bias = use_bias

    def build(self, input_shape):
        self.kernel = self.add_weight("kernel", [int(input_shape[-1]), self.units])
        if self.use_bias:
            self.bias = self.add_weight("bias", [self.units], initializer="zeros")

    def call(self, x):
        x = np.asarray(x)
        k = np.asarray(self.kernel)
        # fp32 sequential-k accumulation (row-vector axpy order), independent of the BLAS in use
        out = np.zeros((x.shape[0], k.shape[1]), dtype=np.float32)
        for j in range(k.shape[0]):
            out += x[:, j:j + 1] * k[j:j + 1, :]
        if self.use_bias:
            out = out + np.asarray(self.bias)
        return _wrap(out)


def _function(f=None, **k):
    if f is None:
        return lambda g: g
    return f


def install():
    tf = types.ModuleType("tensorflow")
    tf.SparseTensor = SparseTensor
    tf.function = _function
    tf.float32 = np.float32
    tf.bool = np.bool_
    tf.constant = lambda v, dtype=None: _wrap(np.array(v, dtype=dtype))
    tf.stack = lambda xs, axis=0: _wrap(np.stack([np.asarray(x) for x in xs], axis=axis))
    tf.concat = lambda xs, axis=0: _wrap(np.concatenate([np.asarray(x) for x in xs], axis=axis))
    tf.split = lambda x, sizes, axis=0: [_wrap(p) for p in np.split(np.asarray(x), np.cumsum(sizes)[:-1], axis=axis)]
    tf.reduce_sum = lambda x, axis=None: _wrap(np.sum(np.asarray(x), axis=axis))
    tf.cast = lambda x, dtype=None: _wrap(np.asarray(x).astype(dtype))
    tf.floor = lambda x: _wrap(np.floor(np.asarray(x)))
    tf.is_tensor = lambda x: isinstance(x, (EagerTensor, SparseTensor))
    tf.stop_gradient = lambda x: x
    tf.zeros_initializer = "zeros"

    sparse = types.ModuleType("tensorflow.sparse")
    sparse.SparseTensor = SparseTensor
    sparse.reorder = _reorder
    sparse.sparse_dense_matmul = _sparse_dense_matmul
    sparse.to_dense = _to_dense
    sparse.retain = _retain
    tf.sparse = sparse

    config = types.SimpleNamespace(experimental=types.SimpleNamespace(
        list_physical_devices=lambda kind=None: [],
        list_logical_devices=lambda kind=None: [],
        set_memory_growth=lambda *a: None))
    config.experimental_run_functions_eagerly = lambda flag: None
    tf.config = config
    tf.random = types.SimpleNamespace(
        uniform=lambda shape: _wrap(_rng.random(tuple(shape)).astype(np.float32)),
        set_seed=lambda s: seed(s))

    keras = types.ModuleType("tensorflow.keras")
    keras.Model = Model
    keras.models = types.SimpleNamespace(Model=Model)
    keras.optimizers = types.SimpleNamespace(get=lambda n: None)
    keras.layers = types.SimpleNamespace(Layer=Layer, ReLU=ReLU, Flatten=Flatten, Dropout=Dropout, Dense=Dense)
    keras.regularizers = types.SimpleNamespace(l2=lambda w: ("l2", w))
    tf.keras = keras
    tf.train = types.SimpleNamespace(Checkpoint=lambda **k: types.SimpleNamespace(**k))
    tf.nn = types.SimpleNamespace()

    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.sparse"] = sparse
    sys.modules["tensorflow.keras"] = keras
    return tf

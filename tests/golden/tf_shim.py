"""A numpy stand-in for the ~30 TensorFlow 1.x symbols the reference's hot-path
modules touch, so that the reference's OWN python (read from /root/reference at
fixture-generation time, never copied) can be executed eagerly in this
container, where TensorFlow cannot be installed.

Each op restates the documented TF semantics in numpy fp32.  The one op whose
values cannot be restated is tf.random_shuffle (TF's Philox stream depends on
graph construction state): here it permutes dim 0 with the oracle's documented
Philox/Fisher-Yates contract (oracle/sampler.py) using (SHUFFLE_SEED,
SHUFFLE_COUNTER), incrementing the counter per call.

Used ONLY by tests/golden/make_golden.py.  Not product code, not oracle code.
"""
import contextlib
import sys
import types

import numpy as np

SHUFFLE_SEED = 123
SHUFFLE_COUNTER = 0
INIT_RNG = np.random.RandomState(2024)


def _arr(x):
    return np.asarray(x)


def install():
    from oracle.sampler import perm_prefix

    tf = types.ModuleType("tensorflow")
    tf.float32, tf.float64, tf.int32, tf.int64 = np.float32, np.float64, np.int32, np.int64

    # ---- flags
    class _Flags(object):
        weight_decay = 0.0
        learning_rate = 0.01
        neg_sample_size = 20
        batch_size = 512

    flags = types.SimpleNamespace(FLAGS=_Flags())
    for n in ("DEFINE_boolean", "DEFINE_string", "DEFINE_float", "DEFINE_integer"):
        setattr(flags, n, lambda *a, **k: None)
    tf.app = types.SimpleNamespace(flags=flags, run=lambda *a, **k: None)

    # ---- scopes / summaries
    @contextlib.contextmanager
    def _scope(*a, **k):
        yield

    tf.variable_scope = _scope
    tf.name_scope = _scope
    tf.summary = types.SimpleNamespace(histogram=lambda *a, **k: None, scalar=lambda *a, **k: None)

    # ---- variables / initialisers
    tf.Variable = lambda initial, name=None, trainable=True, **k: _arr(initial)
    tf.constant = lambda v, dtype=None, **k: _arr(v).astype(dtype) if dtype else _arr(v)
    tf.zeros = lambda shape, dtype=np.float32, **k: np.zeros(shape, dtype=dtype)
    tf.ones = lambda shape, dtype=np.float32, **k: np.ones(shape, dtype=dtype)

    def random_uniform(shape, minval=0.0, maxval=1.0, dtype=np.float32, **k):
        return INIT_RNG.uniform(minval, maxval, size=shape).astype(dtype)

    tf.random_uniform = random_uniform

    def xavier_initializer(**k):
        def init(shape):
            r = np.sqrt(6.0 / (shape[0] + shape[1]))
            return INIT_RNG.uniform(-r, r, size=shape).astype(np.float32)
        return init

    tf.contrib = types.SimpleNamespace(layers=types.SimpleNamespace(
        xavier_initializer=xavier_initializer, l2_regularizer=lambda s: None))

    def get_variable(name, shape=None, dtype=np.float32, initializer=None, regularizer=None, **k):
        if initializer is None:
            initializer = xavier_initializer()
        return initializer(tuple(shape))

    tf.get_variable = get_variable
    tf.train = types.SimpleNamespace(AdamOptimizer=lambda **k: None)

    # ---- array ops
    tf.transpose = lambda x, perm=None: np.transpose(_arr(x), perm)
    tf.reshape = lambda x, shape: np.reshape(_arr(x), tuple(int(s) for s in shape))
    tf.shape = lambda x: _arr(x).shape
    tf.concat = lambda values=None, axis=None, **k: np.concatenate([_arr(v) for v in values], axis=axis)
    tf.expand_dims = lambda x, axis: np.expand_dims(_arr(x), axis)
    tf.cast = lambda x, dtype: _arr(x).astype(dtype)

    def slice_(x, begin, size):
        x = _arr(x)
        idx = tuple(slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
        return x[idx]

    tf.slice = slice_

    def random_shuffle(x, seed=None, name=None):
        global SHUFFLE_COUNTER
        x = _arr(x)
        perm = perm_prefix(SHUFFLE_SEED, SHUFFLE_COUNTER, x.shape[0], x.shape[0])
        SHUFFLE_COUNTER += 1
        return x[perm]

    tf.random_shuffle = random_shuffle

    # ---- math
    tf.matmul = lambda a, b: _arr(a) @ _arr(b)
    tf.add_n = lambda vals: sum(_arr(v) for v in vals[1:]) + _arr(vals[0])
    tf.reduce_mean = lambda x, axis=None: _arr(x).mean(axis=axis, dtype=_arr(x).dtype)
    tf.reduce_max = lambda x, axis=None: _arr(x).max(axis=axis)
    tf.reduce_sum = lambda x, axis=None: _arr(x).sum(axis=axis, dtype=_arr(x).dtype)

    def dropout(x, keep_prob, **k):
        assert float(keep_prob) == 1.0, "golden vectors are generated at dropout=0 only"
        return _arr(x)

    def l2_normalize(x, dim, epsilon=1e-12):
        x = _arr(x)
        ss = (x * x).sum(axis=dim, keepdims=True, dtype=x.dtype)
        return x / np.sqrt(np.maximum(ss, epsilon))

    tf.nn = types.SimpleNamespace(
        embedding_lookup=lambda params, ids: _arr(params)[_arr(ids).astype(np.int64)],
        relu=lambda x, name=None: np.maximum(_arr(x), 0),
        dropout=dropout,
        l2_normalize=l2_normalize,
        sigmoid=lambda x: 1.0 / (1.0 + np.exp(-_arr(x))),
        tanh=lambda x: np.tanh(_arr(x)),
    )
    # ---- loss-head ops (reference prediction.py, models.py:384-405, supervised_models.py:101-126); formulas as
    # documented for TensorFlow: sigmoid xent = max(x, 0) - x z + log(1 + exp(-|x|)); softmax xent = -sum z log_softmax(x);
    # l2_loss = sum(t^2) / 2; top_k sorts descending and breaks ties by the lower index first
    def sigmoid_xent(labels=None, logits=None, **k):
        x, z = _arr(logits), _arr(labels)
        return np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))

    def _log_softmax(x):
        x = _arr(x)
        m = x.max(axis=-1, keepdims=True)
        return x - m - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))

    def softmax_xent(labels=None, logits=None, **k):
        return -(_arr(labels) * _log_softmax(logits)).sum(axis=-1)

    def top_k(x, k=1, **kw):
        x = _arr(x)
        idx = np.argsort(-x, axis=-1, kind="stable")[..., :int(k)]
        return np.take_along_axis(x, idx, axis=-1), idx.astype(np.int32)

    tf.nn.sigmoid_cross_entropy_with_logits = sigmoid_xent
    tf.nn.softmax_cross_entropy_with_logits = softmax_xent
    tf.nn.softmax = lambda x: np.exp(_log_softmax(x))
    tf.nn.l2_loss = lambda t: (_arr(t) * _arr(t)).sum(dtype=_arr(t).dtype) / 2
    tf.nn.top_k = top_k
    tf.ones_like = lambda x: np.ones_like(_arr(x))
    tf.zeros_like = lambda x: np.zeros_like(_arr(x))
    tf.div = lambda a, b: _arr(a) / _arr(b)
    tf.log = lambda x: np.log(_arr(x))
    tf.exp = lambda x: np.exp(_arr(x))
    tf.subtract = lambda a, b, name=None: _arr(a) - _arr(b)

    sys.modules["tensorflow"] = tf
    return tf

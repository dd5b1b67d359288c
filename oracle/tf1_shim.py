"""TEST INFRASTRUCTURE -- a torch-backed, eager stand-in for the part of the TensorFlow-1.8 API that the reference's
``TLSAN/model.py`` uses, so that the UNMODIFIED reference file can be imported and executed in a container without
TensorFlow (``oracle/make_model_golden.py``).  Only ``tests/`` and the golden-vector generator may use it.

How it works: ``install(feeds, variables, dtype)`` puts this module into ``sys.modules['tensorflow']`` and binds

* ``tf.placeholder`` calls, in creation order (model.py:27-53: u, u_cate, i, y, hist_i, hist_i_new, hist_t, sl, sl_new,
  lr, is_training), to the values of ``feeds`` -- the graph is therefore *executed while it is built*: constructing
  ``Model(config, item_cate_list)`` runs build_model and init_optimizer on the bound batch;
* ``tf.get_variable`` / ``tf.layers.dense`` to ``variables[scoped_name]`` (the TF variable names of the checkpoint,
  e.g. ``all/long_term/num_blocks0_0/dense/kernel``), as torch leaf tensors with ``requires_grad``;
* ``tf.gradients`` to ``torch.autograd.grad`` (dense gradients: duplicate indices of a gather are summed, as
  ``apply_gradients`` does before the update), ``tf.clip_by_global_norm`` to the dense-gradient reading
  (``t * clip / max(norm, clip)``), ``GradientDescentOptimizer.apply_gradients`` to ``w - lr * g``.

Every op below restates the documented TF-1.8 semantics of ONE public API entry; nothing here knows about TLSAN.
``gradients`` additionally evaluates the global norm the way TF 1.8 does on the output of tf.gradients (one
IndexedSlices per lookup into a variable, concatenated un-aggregated, ``global_norm`` squares ``.values``) over the
lookups the executed graph really made -- a restatement of TF internals from its sources, applied generically.
What this does NOT reproduce: TF's kernels' floating-point summation order (the run is done in float64 to take
rounding out of the comparison) -- see DESIGN.md section 2.
"""
import contextlib
import sys
import types

import numpy as np
import torch

_S = types.SimpleNamespace(feeds=None, feed_i=0, variables=None, dtype=torch.float64, scope=[], trainable=[],
                           created={}, last_grads=None, last_norm=None, last_norm_tf=None, applied=None, slices=[])

float32 = "float32"
int32 = "int32"
int64 = "int64"
bool = "bool"          # noqa: A001  (tf.bool)


def _dt(d):
    return {float32: _S.dtype, int32: torch.int64, int64: torch.int64, bool: torch.bool}.get(d, d)


class _Shape(tuple):
    def as_list(self):
        return list(self)


# tensors are plain torch tensors; the reference calls x.get_shape().as_list() / x.get_shape()[k]
torch.Tensor.get_shape = lambda self: _Shape(int(d) for d in self.shape)


def install(feeds, variables, dtype=torch.float64):
    """feeds: list of values in placeholder creation order; variables: {tf variable name: numpy array}."""
    _S.feeds, _S.feed_i, _S.dtype = list(feeds), 0, dtype
    _S.variables = {k: np.asarray(v) for k, v in variables.items()}
    _S.scope, _S.trainable, _S.created, _S.slices = [], [], {}, []
    _S.last_grads = _S.last_norm = _S.last_norm_tf = _S.applied = None
    sys.modules["tensorflow"] = sys.modules[__name__]
    return sys.modules[__name__]


def state():
    return _S


# ----------------------------------------------------------------------------------------------- graph plumbing
def placeholder(dtype, shape=None, name=None):
    v = _S.feeds[_S.feed_i]
    _S.feed_i += 1
    if dtype == bool:
        return builtins_bool(v)
    return torch.as_tensor(np.asarray(v), dtype=_dt(dtype))


def builtins_bool(v):
    return True if v else False


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _S.scope.append(name)
    try:
        yield types.SimpleNamespace(name="/".join(_S.scope))
    finally:
        _S.scope.pop()


def get_variable_scope():
    return types.SimpleNamespace(name="/".join(_S.scope))


def constant_initializer(value):
    return ("const", float(value))


def get_variable(name, shape=None, dtype=None, initializer=None):
    full = "/".join(_S.scope + [name])
    if full in _S.created:
        return _S.created[full]
    if full not in _S.variables:
        raise KeyError("the reference asked for variable %r, which the caller did not provide" % full)
    val = _S.variables[full]
    want = tuple(int(d) for d in (shape if shape is not None else val.shape))
    assert tuple(val.shape) == want, (full, val.shape, want)
    if initializer is not None and initializer[0] == "const":
        pass            # the caller's value replaces the initial value; the shape check above is what matters
    t = torch.tensor(val, dtype=_S.dtype, requires_grad=True)
    t.tf_name = full
    _S.created[full] = t
    _S.trainable.append(t)
    return t


class Variable(object):          # tf.Variable(0, trainable=False, name=...)
    def __init__(self, value, trainable=True, name=None):
        assert not trainable
        self.value, self.name = value, name

    def __add__(self, other):
        return self.value + other

    def eval(self, session=None):
        return self.value


def assign(var, value):
    return ("assign", var, value)


def trainable_variables():
    return list(_S.trainable)


class GraphKeys:
    TRAINABLE_VARIABLES = "trainable_variables"


def get_collection(key, scope=None):
    return [v for v in _S.trainable if scope is None or v.tf_name.startswith(scope)]


def add_to_collection(name, value):
    pass


# ----------------------------------------------------------------------------------------------- ops (tf.*)
def convert_to_tensor(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x)


def identity(x):
    return x


def cast(x, dtype):
    return x.to(_dt(dtype))


def shape(x):
    return [int(d) for d in x.shape]


def concat(values, axis):
    return torch.cat(list(values), dim=axis)


def split(value, num_or_size_splits, axis=0):
    assert value.shape[axis] % num_or_size_splits == 0
    return list(torch.split(value, value.shape[axis] // num_or_size_splits, dim=axis))


def reshape(x, shape):
    return x.reshape([int(d) for d in shape])


def squeeze(x, axis=None):
    for a in sorted(axis or [], reverse=True):
        x = x.squeeze(a)
    return x


def expand_dims(x, axis):
    return x.unsqueeze(axis)


def tile(x, multiples):
    return x.repeat(*[int(m) for m in multiples])


def _lookup(params, ids):
    """params[ids] along axis 0.  A lookup into a trainable variable is remembered: its gradient is ONE IndexedSlices
    of tf.gradients (values = the gradient of this output, indices = ids), see gradients() below."""
    out = params[ids]
    if isinstance(params, torch.Tensor) and hasattr(params, "tf_name"):
        _S.slices.append((params.tf_name, ids, out))
    return out


def gather(params, indices):
    """tf.gather along axis 0; params may be a python list (item_cate_list)."""
    p = params if isinstance(params, torch.Tensor) else torch.as_tensor(np.asarray(params), dtype=torch.int64)
    return _lookup(p, indices)


def multiply(a, b):
    return a * b


def add(a, b, name=None):
    return a + b


def add_n(xs):
    out = xs[0]
    for x in xs[1:]:
        out = out + x
    return out


def matmul(a, b, transpose_b=False):
    return a @ (b.transpose(-1, -2) if transpose_b else b)


def reduce_sum(x, axis=None):
    return x.sum() if axis is None else x.sum(dim=axis)


def reduce_mean(x, axis=None):
    return x.mean() if axis is None else x.mean(dim=axis)


def sequence_mask(lengths, maxlen):
    """mask[b, t] = t < lengths[b]"""
    return torch.arange(int(maxlen))[None, :] < lengths[:, None]


def cond(pred, true_fn, false_fn):
    return true_fn() if pred else false_fn()


class _NN:
    @staticmethod
    def embedding_lookup(params, ids):
        return _lookup(params, ids)

    @staticmethod
    def l2_loss(t):
        return (t * t).sum() / 2

    @staticmethod
    def softmax(logits, axis=-1):
        return torch.softmax(logits, dim=axis)

    @staticmethod
    def relu(x):
        return torch.relu(x)

    @staticmethod
    def elu(x):
        return torch.nn.functional.elu(x)

    @staticmethod
    def dropout(x, keep_prob):
        raise NotImplementedError("dropout > 0 is outside the path (train.py default 0)")

    @staticmethod
    def sigmoid_cross_entropy_with_logits(logits=None, labels=None):
        """max(x, 0) - x * z + log(1 + exp(-|x|))"""
        x, z = logits, labels
        return torch.clamp(x, min=0) - x * z + torch.log1p(torch.exp(-torch.abs(x)))


nn = _NN()


class _Layers:
    @staticmethod
    def dense(inputs, units):
        """tf.layers.dense without activation: variables <scope>/dense/kernel [in, units], <scope>/dense/bias [units]"""
        with variable_scope("dense"):
            k = get_variable("kernel", [inputs.shape[-1], units])
            b = get_variable("bias", [units])
        return inputs @ k + b


layers = _Layers()


class _Contrib:
    class layers:          # noqa: N801
        @staticmethod
        def batch_norm(*a, **k):
            raise NotImplementedError("enable_bn is False on the path (model.py:378-381)")


contrib = _Contrib()


# ----------------------------------------------------------------------------------------------- gradients / optimizer
def gradients(ys, xs):
    """Dense gradients (what apply_gradients ends up applying), plus the global norm AS TF 1.8 COMPUTES IT on the
    output of tf.gradients: every lookup into a variable contributes one IndexedSlices whose `.values` are the
    gradient of that lookup's output; gradients_impl._AggregatedGrads CONCATENATES the slices of a variable
    (a dense contribution, e.g. d l2_loss, becomes one more slice over all rows) without summing duplicate rows, and
    clip_ops.global_norm squares `.values` -- so norm_tf^2 = sum over variables of
    (sum over its lookups |d out|^2  +  |dense remainder|^2).  Restated from the TF-1.8 sources; the lookups are the
    ones the reference graph really makes."""
    outs = [o for (_, _, o) in _S.slices]
    both = torch.autograd.grad(ys, list(xs) + outs, allow_unused=True)
    g = [torch.zeros_like(x) if gi is None else gi for gi, x in zip(both[:len(xs)], xs)]
    _S.last_grads = {x.tf_name: gi.detach().numpy().copy() for gi, x in zip(g, xs)}
    sq = 0.0
    for x, gi in zip(xs, g):
        rest = gi.detach().clone()
        for (name, ids, o), go in zip(_S.slices, both[len(xs):]):
            if name != x.tf_name or go is None:
                continue
            sq += float((go * go).sum())
            idx = torch.as_tensor(np.asarray(ids)).reshape(-1) if not isinstance(ids, torch.Tensor) else ids.reshape(-1)
            rest.index_add_(0, idx, -go.detach().reshape((idx.numel(),) + tuple(rest.shape[1:])))
        sq += float((rest * rest).sum())
    _S.last_norm_tf = sq ** 0.5
    return list(g)


def clip_by_global_norm(t_list, clip_norm):
    norm = torch.sqrt(sum((t * t).sum() for t in t_list))
    _S.last_norm = float(norm)
    scale = clip_norm / max(float(norm), clip_norm)
    return [t * scale for t in t_list], norm


class _Optimizer(object):
    def __init__(self, learning_rate=None):
        self.lr = learning_rate


class _SGD(_Optimizer):
    def apply_gradients(self, grads_and_vars, global_step=None):
        lr = float(self.lr)
        _S.applied = {v.tf_name: (v.detach() - lr * g.detach()).numpy().copy() for g, v in grads_and_vars}
        return ("train_op",)


class _Unsupported(_Optimizer):
    def apply_gradients(self, *a, **k):
        raise NotImplementedError("the shim pins the default optimizer (sgd) only")


class _Saver(object):
    def save(self, *a, **k):
        raise NotImplementedError

    def restore(self, *a, **k):
        raise NotImplementedError


train = types.SimpleNamespace(GradientDescentOptimizer=_SGD, AdamOptimizer=_Unsupported, AdadeltaOptimizer=_Unsupported,
                              RMSPropOptimizer=_Unsupported, Saver=_Saver)


# ----------------------------------------------------------------------------------------------- summaries / metrics
class _Summary:
    class FileWriter(object):
        def __init__(self, path):
            self.path = path

        def add_summary(self, *a, **k):
            pass

    @staticmethod
    def histogram(name, value):
        return ("histogram", name)

    @staticmethod
    def scalar(name, value):
        return ("scalar", name)

    @staticmethod
    def merge(xs):
        return ("merge", xs)


summary = _Summary()


class _Metrics:
    """tf.metrics.precision_at_k / recall_at_k for single-label rows: one batch's value (the streaming accumulation of
    the real op is a running mean of exactly these counts, restated in the oracle's StreamingTopK)."""

    @staticmethod
    def _hits(labels, predictions, k):
        top = torch.topk(predictions, int(k), dim=1).indices
        return (top == labels[:, None]).any(dim=1).to(predictions.dtype).sum()

    @staticmethod
    def precision_at_k(labels=None, predictions=None, k=1):
        v = _Metrics._hits(labels, predictions, k) / (predictions.shape[0] * k)
        return v, v

    @staticmethod
    def recall_at_k(labels=None, predictions=None, k=1):
        v = _Metrics._hits(labels, predictions, k) / predictions.shape[0]
        return v, v


metrics = _Metrics()

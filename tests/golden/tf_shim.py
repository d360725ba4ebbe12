"""A minimal TensorFlow-1.x op shim over torch CPU tensors.

Used ONLY by tests/golden/make_golden.py to execute the reference's own Python
source (nets/model.py, nets/pixellink.py, nets/model_vgg_16.py) verbatim, since
TensorFlow 1.4 cannot be installed offline.  Every op below
restates the documented TF semantics eagerly; gradients come from torch autograd,
which matches TF autodiff for these ops (softmax, sparse xent, reductions;
comparisons / casts of masks carry no gradient in either).
"""
from __future__ import annotations

import contextlib
import types

import torch

float32 = torch.float32
int32 = torch.int32
uint8 = torch.uint8
bool_ = torch.bool


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(x, dtype=dtype)


def shape(x):
    return list(_t(x).shape)


def cast(x, dtype):
    x = _t(x)
    if dtype in (torch.int32, torch.int64) and x.is_floating_point():
        return torch.trunc(x).to(dtype)  # tf.cast float->int truncates toward zero
    return x.to(dtype)


def reshape(x, shp):
    return _t(x).reshape([int(s) for s in shp])


def equal(a, b):
    return _t(a) == b


def logical_and(a, b):
    return torch.logical_and(_t(a), _t(b))


def logical_not(a):
    return torch.logical_not(_t(a))


def reduce_sum(x, axis=None):
    if isinstance(x, (list, tuple)):
        x = torch.stack([_t(v) for v in x])
    x = _t(x)
    if x.dtype == torch.bool:
        x = x.to(torch.int32)
    return x.sum() if axis is None else x.sum(dim=axis)


def reduce_mean(x, axis=None):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=axis)


def minimum(a, b):
    return torch.minimum(_t(a), _t(b))


def maximum(a, b):
    return torch.maximum(_t(a), _t(b))


def boolean_mask(x, mask):
    return _t(x)[_t(mask)]


def where(c, a, b):
    return torch.where(_t(c), _t(a), _t(b))


def zeros_like(x, dtype=None):
    return torch.zeros_like(_t(x), dtype=dtype)


def stack(xs, axis=0):
    return torch.stack(list(xs), dim=axis)


def split(value, num_or_size_splits, axis):
    return torch.chunk(value, num_or_size_splits, dim=axis)


def add_n(xs):
    out = xs[0]
    for v in xs[1:]:
        out = out + v
    return out


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def constant(v, dtype=None):
    return torch.tensor(v, dtype=dtype or torch.float32)


def cond(pred, true_fn, false_fn):
    return true_fn() if bool(pred) else false_fn()


@contextlib.contextmanager
def name_scope(_name):
    yield


class _NN(types.SimpleNamespace):
    @staticmethod
    def top_k(x, k):
        k = int(k)
        vals, idx = torch.topk(_t(x), k, largest=True, sorted=True)
        return vals, idx

    @staticmethod
    def softmax(x):
        return torch.softmax(_t(x), dim=-1)

    @staticmethod
    def sparse_softmax_cross_entropy_with_logits(logits=None, labels=None):
        logits = _t(logits)
        labels = _t(labels).to(torch.int64)
        lsm = torch.log_softmax(logits, dim=-1)
        return -torch.gather(lsm, -1, labels.unsqueeze(-1)).squeeze(-1)


nn = _NN()


class _Summary(types.SimpleNamespace):
    @staticmethod
    def scalar(*a, **k):
        return None

    @staticmethod
    def image(*a, **k):
        return None

    @staticmethod
    def histogram(*a, **k):
        return None


summary = _Summary()


class GraphKeys:
    LOSSES = "losses"


_collections = {}


def add_to_collection(name, value):
    _collections.setdefault(name, []).append(value)


def get_collection(name):
    return list(_collections.get(name, []))


def reset_collections():
    _collections.clear()


class _Slim(types.SimpleNamespace):
    @staticmethod
    def softmax(x):
        return torch.softmax(_t(x), dim=-1)


slim = _Slim()

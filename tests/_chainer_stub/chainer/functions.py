"""numpy implementations of the chainer.functions used by the hot path."""
import numpy as np

from .variable import Variable, _unwrap


def _v(x):
    return Variable(x)


def matmul(a, b, transa=False, transb=False):
    a, b = _unwrap(a), _unwrap(b)
    if transa:
        a = np.swapaxes(a, -1, -2)
    if transb:
        b = np.swapaxes(b, -1, -2)
    return _v(np.matmul(a, b))


def transpose(x, axes=None):
    return _v(np.transpose(_unwrap(x), axes))


def squeeze(x, axis=None):
    return _v(np.squeeze(_unwrap(x), axis=axis))


def expand_dims(x, axis):
    return _v(np.expand_dims(_unwrap(x), axis))


def repeat(x, repeats, axis=None):
    return _v(np.repeat(_unwrap(x), repeats, axis=axis))


def cast(x, typ):
    return _v(_unwrap(x).astype(typ))


def stack(xs, axis=0):
    return _v(np.stack([_unwrap(x) for x in xs], axis=axis))


def concat(xs, axis=1):
    return _v(np.concatenate([_unwrap(x) for x in xs], axis=axis))


def where(cond, a, b):
    return _v(np.where(_unwrap(cond), _unwrap(a), _unwrap(b)))


def sum(x, axis=None, keepdims=False):
    return _v(np.sum(_unwrap(x), axis=axis, keepdims=keepdims))


def mean(x, axis=None, keepdims=False):
    return _v(np.mean(_unwrap(x), axis=axis, keepdims=keepdims))


def minimum(a, b):
    return _v(np.minimum(_unwrap(a), _unwrap(b)))


def maximum(a, b):
    return _v(np.maximum(_unwrap(a), _unwrap(b)))


def batch_inv(a):
    return _v(np.linalg.inv(_unwrap(a)))


def split_axis(x, indices_or_sections, axis, force_tuple=True):
    return tuple(_v(p) for p in np.split(_unwrap(x), indices_or_sections, axis=axis))


def separate(x, axis=0):
    x = _unwrap(x)
    return tuple(_v(np.take(x, i, axis=axis)) for i in range(x.shape[axis]))


def sqrt(x): return _v(np.sqrt(_unwrap(x)))
def sin(x): return _v(np.sin(_unwrap(x)))
def cos(x): return _v(np.cos(_unwrap(x)))
def exp(x): return _v(np.exp(_unwrap(x)))
def arctan2(a, b): return _v(np.arctan2(_unwrap(a), _unwrap(b)))
def clip(x, lo, hi): return _v(np.clip(_unwrap(x), lo, hi))
def sigmoid(x): return _v(1.0 / (1.0 + np.exp(-_unwrap(x))))
def reshape(x, shape): return _v(np.reshape(_unwrap(x), shape))
def absolute(x): return _v(np.abs(_unwrap(x)))
def square(x): return _v(np.square(_unwrap(x)))
def diagonal(x, *a, **k): return _v(np.diagonal(_unwrap(x), *a, **k))

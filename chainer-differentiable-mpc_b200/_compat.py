"""Chainer-optional glue shared by the facade modules.

The reference's operators are Chainer FunctionNodes / Links.  Chainer is imported
lazily: when it is importable (real Chainer on a user's box, or the forward-only
stub under tests/_chainer_stub) the facade classes derive from its base classes and
return `chainer.Variable`s, exactly like the reference; when it is not, small
stand-ins with the same call protocol are used and plain numpy arrays are returned.
"""
import numpy as np

try:  # pragma: no cover - depends on the environment
    import chainer as _chainer
    from chainer import function_node as _fn
    HAVE_CHAINER = True
except Exception:  # ImportError or a broken install
    _chainer = None
    _fn = None
    HAVE_CHAINER = False


def to_xp(x):
    """util.to_xp (util.py:61-64): unwrap a Variable, pass arrays / None through."""
    if x is None:
        return None
    if HAVE_CHAINER and isinstance(x, _chainer.Variable):
        return x.array
    if hasattr(x, "array") and not isinstance(x, np.ndarray):
        return x.array
    return x


def wrap(x):
    """Return x the way the reference would hand it back (Variable when Chainer is present)."""
    if HAVE_CHAINER and x is not None and not isinstance(x, _chainer.Variable):
        return _chainer.Variable(x)
    return x


class _PlainFunctionNode:
    """Minimal FunctionNode protocol (apply -> forward on raw arrays; retain_*)."""

    def __init__(self):
        self._in = None
        self._out = None
        self._ri = ()
        self._ro = ()

    def retain_inputs(self, idx):
        self._ri = tuple(idx)

    def retain_outputs(self, idx):
        self._ro = tuple(idx)

    def apply(self, inputs):
        self._in = tuple(to_xp(v) for v in inputs)
        out = self.forward(self._in)
        if not isinstance(out, tuple):
            out = (out,)
        self._out = tuple(to_xp(v) for v in out)
        return self._out

    def get_retained_inputs(self):
        return tuple(self._in[i] for i in self._ri)

    def get_retained_outputs(self):
        return tuple(self._out[i] for i in self._ro)


class _PlainParameter:
    """chainer.Parameter stand-in when Chainer is absent: `.array`, `.grad`, numpy conversion."""

    def __init__(self, initializer=None, name=None):
        self.array = None if initializer is None else np.array(initializer)
        self.grad = None
        self.name = name

    data = property(lambda self: self.array)
    shape = property(lambda self: self.array.shape)
    dtype = property(lambda self: self.array.dtype)

    def __array__(self, dtype=None, copy=None):
        return self.array if dtype is None else self.array.astype(dtype)

    def cleargrad(self):
        self.grad = None


class _PlainLink:
    """chainer.Link stand-in: init_scope, __call__ -> forward, params(), cleargrads()."""
    xp = np

    def __init__(self):
        pass

    def init_scope(self):
        import contextlib
        return contextlib.nullcontext()

    def __call__(self, *a, **k):
        return self.forward(*a, **k)

    def params(self):
        for v in self.__dict__.values():
            if isinstance(v, _PlainParameter):
                yield v

    def cleargrads(self):
        for p in self.params():
            p.grad = None


FunctionNodeBase = _fn.FunctionNode if HAVE_CHAINER else _PlainFunctionNode
LinkBase = _chainer.Link if HAVE_CHAINER else object
NetLinkBase = _chainer.Link if HAVE_CHAINER else _PlainLink      # base of the LqrNet* / MpcNet_* model links
Parameter = _chainer.Parameter if HAVE_CHAINER else _PlainParameter


def as_f(x, dtype=None):
    """C-contiguous floating array (float64 unless the input is float32)."""
    a = np.asarray(to_xp(x))
    if dtype is None:
        dtype = np.float32 if a.dtype == np.float32 else np.float64
    return np.ascontiguousarray(a, dtype=dtype)

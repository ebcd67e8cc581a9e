import numpy as np


def _unwrap(x):
    return x.array if isinstance(x, Variable) else x


class Variable:
    """numpy-backed value holder with Chainer's surface (no graph)."""
    __array_priority__ = 200  # ndarray (op) Variable defers to Variable.__r*__

    def __init__(self, data=None, name=None, requires_grad=True):
        self.array = _unwrap(data)
        self.grad = None
        self.name = name

    # -- attributes -------------------------------------------------------
    @property
    def data(self):
        return self.array

    @data.setter
    def data(self, v):
        self.array = v

    @property
    def shape(self):
        return self.array.shape

    @property
    def dtype(self):
        return self.array.dtype

    @property
    def ndim(self):
        return self.array.ndim

    @property
    def size(self):
        return self.array.size

    @property
    def T(self):
        return Variable(self.array.T)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return Variable(self.array.reshape(*shape))

    def __len__(self):
        return len(self.array)

    def __getitem__(self, idx):
        return Variable(self.array[idx])

    def __iter__(self):
        for i in range(len(self.array)):
            yield Variable(self.array[i])

    def __repr__(self):
        return "variable(%r)" % (self.array,)

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self.array)
        return a.astype(dtype) if dtype is not None else a

    def cleargrad(self):
        self.grad = None

    def backward(self, *a, **k):
        raise NotImplementedError("chainer stub is forward-only")

    # -- arithmetic -------------------------------------------------------
    def _bin(self, other, op):
        return Variable(op(self.array, _unwrap(other)))

    def _rbin(self, other, op):
        return Variable(op(_unwrap(other), self.array))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._rbin(o, np.add)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._rbin(o, np.subtract)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._rbin(o, np.multiply)
    def __truediv__(self, o): return self._bin(o, np.true_divide)
    def __rtruediv__(self, o): return self._rbin(o, np.true_divide)
    def __matmul__(self, o): return self._bin(o, np.matmul)
    def __rmatmul__(self, o): return self._rbin(o, np.matmul)
    def __pow__(self, o): return self._bin(o, np.power)
    def __rpow__(self, o): return self._rbin(o, np.power)
    def __neg__(self): return Variable(-self.array)
    def __abs__(self): return Variable(np.abs(self.array))

    def __iadd__(self, o):
        self.array = self.array + _unwrap(o)
        return self

    def __isub__(self, o):
        self.array = self.array - _unwrap(o)
        return self

    def __imul__(self, o):
        self.array = self.array * _unwrap(o)
        return self

    # comparisons return raw arrays (as Chainer does not define them on
    # Variable, the hot path only compares .array / to_xp'ed values)
    def __lt__(self, o): return self.array < _unwrap(o)
    def __le__(self, o): return self.array <= _unwrap(o)
    def __gt__(self, o): return self.array > _unwrap(o)
    def __ge__(self, o): return self.array >= _unwrap(o)


class Parameter(Variable):
    def __init__(self, initializer=None, shape=None, name=None):
        super().__init__(np.array(_unwrap(initializer)) if initializer is not None else None, name=name)

    def update(self):
        pass


def as_variable(x):
    return x if isinstance(x, Variable) else Variable(x)

import contextlib

import numpy as _np

from .variable import Parameter


class Link:
    xp = _np

    def __init__(self, **kw):
        pass

    @contextlib.contextmanager
    def init_scope(self):
        yield

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def params(self, include_uninit=True):
        for v in self.__dict__.values():
            if isinstance(v, Parameter):
                yield v
            elif isinstance(v, Link):
                yield from v.params()

    def cleargrads(self):
        for p in self.params():
            p.grad = None


class Chain(Link):
    pass

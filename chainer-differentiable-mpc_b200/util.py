"""Host-side helpers of the B200 build with the reference's names (reference util.py).

Only what callers of the hot path import is provided: QuadCost / LinDx (util.py:25-32),
to_xp (:61-64), expand_time_batch (:361-377), the xp* batched helpers used by user code
around the solver (:117-123, :293-358, :427-434) and get_traj / get_cost (:126-236) for
LinDx / callable dynamics.  The batched-LU wrappers (util.py:437-528, torch.lu /
torch.lu_solve / scipy) are intentionally absent: factorisation happens inside the CUDA
kernels and torch is no longer imported by the hot path.
"""
from collections import namedtuple

import numpy as np

from _compat import HAVE_CHAINER, to_xp, wrap  # noqa: F401

QuadCost = namedtuple("QuadCost", "C c")
LinDx = namedtuple("LinDx", "F f")
QuadCost.__new__.__defaults__ = (None,) * len(QuadCost._fields)
LinDx.__new__.__defaults__ = (None,) * len(LinDx._fields)


def get_array_module(a):
    return np


def xpclamp(x, lower, upper):
    assert x.shape == lower.shape and x.shape == upper.shape
    assert (lower <= upper).all()
    return np.minimum(np.maximum(x, lower), upper)


def xpbmv(a, x):
    assert a.shape[0] == x.shape[0] and a.shape[2] == x.shape[1] and x.ndim == 2
    return np.einsum("bij,bj->bi", a, x)


def xpbger(x, y):
    return np.einsum("bi,bj->bij", x, y)


def xpbquad(x, Q):
    return np.einsum("bi,bij,bj->b", x, Q, x)


def xpbdot(x, y):
    return np.einsum("bi,bi->b", x, y)


def xpexpand_batch(m, n_batch):
    return np.repeat(np.expand_dims(m, 0), n_batch, axis=0)


def expand_time_batch(m, time, n_batch):
    """[...] -> [time, n_batch, ...] (util.py:361-377).  With Chainer present the broadcast is
    done with chainer.functions so gradients flow back (sum over T,B) as in the reference."""
    if HAVE_CHAINER:
        from chainer import functions as F
        v = F.expand_dims(F.expand_dims(m, 0), 0)
        v = F.repeat(v, n_batch, axis=1)
        return F.repeat(v, time, axis=0)
    a = np.asarray(to_xp(m))
    return np.broadcast_to(a, (time, n_batch) + a.shape).copy()

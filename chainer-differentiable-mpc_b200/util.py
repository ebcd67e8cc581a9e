"""Host-side helpers of the B200 build under the reference's names (reference util.py:25-434).

Every helper a caller of the hot path imports from the reference's `util` is here, so that the reference's own
callers (`env_dx/il_env.py:20`, `mpc/mpc_net.py:18`, `experiment_mpc/MpcNet.py:36`, `examples/*.py`) import against this
package unchanged:

  QuadCost, LinDx (:25-32) · chainer_diag (:37-50) · to_xp (:61-64) · table_log (:67-93) · get_array_module (:96-100)
  clamp / xpclamp (:103-123) · get_cost / xpget_cost (:126-198) · get_traj / xpget_traj (:201-277)
  bmv / xpbmv (:280-298) · bger / xpbger (:301-329) · bquad / xpbquad (:332-358) · expand_time_batch (:361-377)
  expand_batch / xpexpand_batch (:380-408) · bdot / xpbdot (:411-434)

The un-prefixed names are the Chainer versions (they take and return `chainer.Variable`s and stay differentiable when
Chainer is importable; without Chainer they fall back to the numpy versions); the `xp*` names work on raw arrays.
`get_traj` / `xpget_traj` / `get_cost` / `xpget_cost` with `LinDx` or pendulum dynamics run on the GPU through
`dmpc_get_traj` (csrc/mpc_kernels.cuh `traj_kernel`); an arbitrary Python callable is rolled out on the host exactly as
the reference does.  The batched-LU wrappers (util.py:437-528: torch.lu / torch.lu_solve / scipy) are intentionally
absent: factorisation happens inside the CUDA kernels and torch is no longer imported by the hot path.
"""
import operator
from collections import namedtuple

import numpy as np

from _compat import HAVE_CHAINER, to_xp, wrap  # noqa: F401

if HAVE_CHAINER:
    from chainer import functions as F
else:
    F = None

QuadCost = namedtuple("QuadCost", "C c")
LinDx = namedtuple("LinDx", "F f")
QuadCost.__new__.__defaults__ = (None,) * len(QuadCost._fields)
LinDx.__new__.__defaults__ = (None,) * len(LinDx._fields)

_seen_tables = []


def get_array_module(a):
    return np


# ---------------------------------------------------------------------------------------- small utilities
def chainer_diag(q):
    """Vector -> diagonal matrix, differentiable w.r.t. q (util.py:37-50)."""
    dim = q.shape[0]
    if HAVE_CHAINER:
        return F.where(np.eye(dim, dtype=bool), q, np.zeros((dim, dim)))      # q broadcasts along rows: (i,i) <- q[i]
    return np.diag(np.asarray(to_xp(q)))


def table_log(tag, d):
    """One table row per call, header on first use of `tag` (util.py:67-93)."""
    def print_row(r):
        print("| " + " | ".join(r) + " |")

    if tag not in _seen_tables:
        print_row(map(operator.itemgetter(0), d))
        _seen_tables.append(tag)
    s = []
    for di in d:
        assert len(di) in [2, 3]
        if len(di) == 3:
            e, fmt = di[1:]
            try:
                s.append(fmt.format(e))
            except Exception:
                s.append(fmt.format(e.data))
        else:
            s.append(str(di[1]))
    print_row(s)


def xpclamp(x, lower, upper):
    assert x.shape == lower.shape, str(x.shape) + " : " + str(lower.shape)
    assert x.shape == upper.shape
    assert (lower <= upper).all()
    return np.minimum(np.maximum(x, lower), upper)


def clamp(x, lower, upper):
    assert x.shape == lower.shape
    assert x.shape == upper.shape
    assert (np.asarray(to_xp(lower)) <= np.asarray(to_xp(upper))).all(), " lower is larger than upper"
    if HAVE_CHAINER:
        return F.minimum(F.maximum(x, lower), upper)
    return xpclamp(np.asarray(to_xp(x)), np.asarray(to_xp(lower)), np.asarray(to_xp(upper)))


# ---------------------------------------------------------------------------------------- batched algebra
def xpbmv(a, x):
    assert a.shape[0] == x.shape[0], "batch mismatch" + str(a.shape) + "," + str(x.shape)
    assert a.shape[2] == x.shape[1], "mat mul dim mismatch"
    assert len(x.shape) == 2, " x is not batch vector"
    return np.squeeze(np.matmul(a, np.expand_dims(x, axis=2)), axis=2)


def bmv(a, x):
    assert a.shape[0] == x.shape[0], "batch mismatch"
    assert a.shape[2] == x.shape[1], "mat mul dim mismatch"
    assert len(x.shape) == 2, " x is not batch vector"
    if HAVE_CHAINER:
        return F.squeeze(F.matmul(a, F.expand_dims(x, axis=2)), axis=2)
    return xpbmv(np.asarray(to_xp(a)), np.asarray(to_xp(x)))


def xpbger(x, y):
    return np.expand_dims(x, 2) @ np.expand_dims(y, 1)


def bger(x, y):
    if HAVE_CHAINER:
        return F.expand_dims(x, 2) @ F.expand_dims(y, 1)
    return xpbger(np.asarray(to_xp(x)), np.asarray(to_xp(y)))


def xpbquad(x, Q):
    assert x.shape[0] == Q.shape[0], "batch mismatch" + str(x.shape) + ":" + str(Q.shape)
    assert x.shape[1] == Q.shape[1], "mat mul dim mismatch"
    assert Q.shape[2] == Q.shape[1], "Q is not square matrix"
    res = np.squeeze(np.squeeze(np.expand_dims(x, 1) @ Q @ np.expand_dims(x, 2), axis=1), axis=1)
    assert list(res.shape) == [list(x.shape)[0]]
    return res


def bquad(x, Q):
    assert x.shape[0] == Q.shape[0], "batch mismatch" + str(x.shape) + ":" + str(Q.shape)
    assert x.shape[1] == Q.shape[1], "mat mul dim mismatch"
    assert Q.shape[2] == Q.shape[1], "Q is not square matrix"
    if HAVE_CHAINER:
        return F.squeeze(F.squeeze(F.expand_dims(x, 1) @ Q @ F.expand_dims(x, 2), axis=1), axis=1)
    return xpbquad(np.asarray(to_xp(x)), np.asarray(to_xp(Q)))


def xpbdot(x, y):
    assert x.shape[0] == y.shape[0]
    assert x.shape[1] == y.shape[1]
    return np.squeeze(np.squeeze(np.expand_dims(x, 1) @ np.expand_dims(y, 2), axis=1), axis=1)


def bdot(x, y):
    assert x.shape[0] == y.shape[0]
    assert x.shape[1] == y.shape[1]
    if HAVE_CHAINER:
        return F.squeeze(F.squeeze(F.expand_dims(x, 1) @ F.expand_dims(y, 2), axis=1), axis=1)
    return xpbdot(np.asarray(to_xp(x)), np.asarray(to_xp(y)))


def xpexpand_batch(m, n_batch):
    m = np.repeat(np.expand_dims(m, 0), n_batch, axis=0)
    assert list(m.shape)[0] == n_batch
    return m


def expand_batch(m, n_batch):
    """[...] -> [n_batch, ...] (util.py:380-392)."""
    if HAVE_CHAINER:
        m = F.repeat(F.expand_dims(m, 0), n_batch, axis=0)
        assert list(m.shape)[0] == n_batch
        return m
    return xpexpand_batch(np.asarray(to_xp(m)), n_batch)


def expand_time_batch(m, time, n_batch):
    """[...] -> [time, n_batch, ...] (util.py:361-377).  With Chainer present the broadcast is done with
    chainer.functions so gradients flow back (sum over T,B) as in the reference; the fused alternative that never
    materialises the [T,B,...] gradient is `dmpc_*_reduced` (DiffLqr/MPCstep.backward_reduced_numpy)."""
    if HAVE_CHAINER:
        v = F.expand_dims(F.expand_dims(m, 0), 0)
        v = F.repeat(v, n_batch, axis=1)
        v = F.repeat(v, time, axis=0)
        assert list(v.shape)[0] == time and list(v.shape)[1] == n_batch
        return v
    a = np.asarray(to_xp(m))
    return np.broadcast_to(a, (time, n_batch) + a.shape).copy()


# ---------------------------------------------------------------------------------------- trajectories and costs
def _device_traj(T, u, x_init, dynamics, device=0):
    """Rollout on the GPU (dmpc_get_traj) for LinDx / pendulum dynamics; None when `dynamics` is another callable."""
    import _native
    u = np.ascontiguousarray(to_xp(u), dtype=np.float64)
    x_init = np.ascontiguousarray(to_xp(x_init), dtype=np.float64)
    Tn, B, m = u.shape
    n = x_init.shape[1]
    assert Tn >= T
    if isinstance(dynamics, LinDx):
        Fm = np.ascontiguousarray(to_xp(dynamics.F), dtype=np.float64)
        f = to_xp(dynamics.f)
        if f is not None:
            f = np.ascontiguousarray(f, dtype=np.float64)
            assert f.shape[1:] == Fm.shape[1:3]
        ctx = _native.default_context(device)
        x = ctx.empty((T, B, n))
        ctx.get_traj(np.float64, T, B, n, m, _native.DYN_LINEAR, ctx.to_device(x_init), ctx.to_device(u[:T]),
                     ctx.to_device(Fm), None if f is None else ctx.to_device(f), None, x)
        return x.download()
    from mpc_step import is_pendulum, pendulum_params       # lazy: mpc_step imports this module
    if is_pendulum(dynamics):
        ctx = _native.default_context(device)
        x = ctx.empty((T, B, n))
        ctx.get_traj(np.float64, T, B, n, m, _native.DYN_PENDULUM, ctx.to_device(x_init), ctx.to_device(u[:T]), None, None,
                     pendulum_params(dynamics), x)
        return x.download()
    return None


def xpget_traj(T, u, x_init, dynamics):
    """State sequence x[T,B,n] of the controls u from x_init (util.py:239-277)."""
    x = _device_traj(T, u, x_init, dynamics)
    if x is not None:
        return x
    xs = [np.asarray(to_xp(x_init))]
    u = np.asarray(to_xp(u))
    for t in range(T - 1):
        xs.append(np.asarray(to_xp(dynamics(xs[t], u[t]))))
    return np.stack(xs, axis=0)


def get_traj(T, u, x_init, dynamics):
    """util.py:201-236.  LinDx / pendulum: one launch of traj_kernel; the result is a constant w.r.t. the Chainer graph
    (the reference's callers use it under no_backprop_mode, box_ddp.py:123)."""
    return wrap(xpget_traj(T, u, x_init, dynamics))


def xpget_cost(T, u, cost, dynamics=None, x_init=None, x=None):
    """Total cost sum_t 0.5 tau^T C_t tau + c_t . tau per batch element (util.py:162-198)."""
    assert x_init is not None or x is not None
    u = np.asarray(to_xp(u))
    if x is None:
        x = xpget_traj(T, u, x_init, dynamics)
    x = np.asarray(to_xp(x))
    objs = []
    for t in range(T):
        xut = np.concatenate((x[t], u[t]), axis=1)
        if isinstance(cost, QuadCost):
            objs.append(0.5 * xpbquad(xut, np.asarray(to_xp(cost.C))[t]) + xpbdot(xut, np.asarray(to_xp(cost.c))[t]))
        else:
            objs.append(np.asarray(to_xp(cost(xut))))
    return np.sum(np.stack(objs, axis=0), axis=0)


def get_cost(T, u, cost, dynamics=None, x_init=None, x=None):
    """util.py:126-159 (Variable in / Variable out)."""
    assert x_init is not None or x is not None
    if not HAVE_CHAINER:
        return xpget_cost(T, u, cost, dynamics, x_init, x)
    if x is None:
        x = get_traj(T, u, x_init, dynamics)
    objs = []
    for t in range(T):
        xut = F.concat((x[t], u[t]))
        if isinstance(cost, QuadCost):
            objs.append(0.5 * bquad(xut, cost.C[t]) + bdot(xut, cost.c[t]))
        else:
            objs.append(cost(xut))
    return F.sum(F.stack(objs, axis=0), axis=0)

"""Taylor models of a Python-callable cost / dynamics - API of reference mpc/approximate.py:18-54, 77-119.

    hessians, grads, costs = approximate_cost(x, u, Cf)         # [T,B,s,s], [T,B,s] (= grad - H tau), [T,B]
    large_F, f = linearize_dynamics(x, u, dynamics)             # [T-1,B,n,s], [T-1,B,n] along the RE-ROLLED trajectory

These are the plugin seams of BoxDDP for costs that are not a util.QuadCost and dynamics that are neither util.LinDx nor
the pendulum (whose step and analytic Jacobian are device code, csrc/mpc_kernels.cuh `pendulum_step`).  A Python
callable cannot run inside a kernel, so this is host code by construction ("the fall-back
is host rollout through the Python callable, one launch for backward_rec only").

Derivatives: the reference differentiates the callable with `chainer.grad` (approximate.py:39, 45, 103).  With a real
Chainer (one that has `chainer.grad`) the same is done here; otherwise - the build image has no Chainer - central
differences in float64, batched over B (2 s evaluations per timestep for a Jacobian, 2 s^2 + 1 for a Hessian), with steps
cbrt(eps) / eps^(1/4) scaled by max(1, |tau_j|): relative error ~1e-10 / ~1e-7 on smooth callables, documented and
tested as such (tests/test_gpu_generic_plugins.py), not bit-identical to autograd.  A kink inside the stencil (a clip in the
dynamics exactly at a clamped control) is detected and resolved to the inclusive one-sided slope (`_fd_column`).
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.dirname(_here)
for _p in (_pkg, _here):
    if _p not in sys.path:
        sys.path.append(_p)

import numpy as np  # noqa: E402

from _compat import HAVE_CHAINER, to_xp, wrap  # noqa: E402

FD_STEP_GRAD = float(np.finfo(np.float64).eps) ** (1.0 / 3.0)
FD_STEP_HESS = float(np.finfo(np.float64).eps) ** 0.25


def _autograd():
    """chainer.grad when a real Chainer provides it (the forward-only test stub does not)."""
    if not HAVE_CHAINER:
        return None
    import chainer
    return getattr(chainer, "grad", None)


def _np(v):
    return np.asarray(to_xp(v), dtype=np.float64)


def _steps(tau, h):
    return h * np.maximum(1.0, np.abs(tau))


# ------------------------------------------------------------------------------------------------ cost
def _cost_taylor_fd(Cf, tau):
    """(H [B,s,s], g [B,s], cost [B]) of Cf at tau [B,s] by central differences; H is symmetrised."""
    B, s = tau.shape
    c0 = _np(Cf(tau))
    assert list(c0.shape) == [B]
    hg = _steps(tau, FD_STEP_GRAD)
    g = np.empty((B, s))
    for j in range(s):
        e = np.zeros((B, s)); e[:, j] = hg[:, j]
        g[:, j] = (_np(Cf(tau + e)) - _np(Cf(tau - e))) / (2.0 * hg[:, j])
    hh = _steps(tau, FD_STEP_HESS)
    H = np.empty((B, s, s))
    for i in range(s):
        ei = np.zeros((B, s)); ei[:, i] = hh[:, i]
        H[:, i, i] = (_np(Cf(tau + ei)) - 2.0 * c0 + _np(Cf(tau - ei))) / hh[:, i] ** 2
        for j in range(i + 1, s):
            ej = np.zeros((B, s)); ej[:, j] = hh[:, j]
            d = (_np(Cf(tau + ei + ej)) - _np(Cf(tau + ei - ej)) - _np(Cf(tau - ei + ej)) + _np(Cf(tau - ei - ej)))
            H[:, i, j] = H[:, j, i] = d / (4.0 * hh[:, i] * hh[:, j])
    return H, g, c0


def _cost_taylor_chainer(Cf, tau, grad):
    import chainer
    from chainer import functions as F
    v = chainer.Variable(tau)
    cost = Cf(v)
    g = grad([F.sum(cost)], [v], enable_double_backprop=True)[0]
    cols = [grad([F.sum(g[:, j])], [v])[0].array for j in range(tau.shape[1])]
    return np.stack(cols, axis=-1), g.array, cost.array


def approximate_cost(x, u, Cf):
    """Quadratic model of the cost callable at every (x_t, u_t) - reference approximate.py:18-54.  `Cf` maps tau [B,s] to
    a cost [B].  Returns (hessians, grads - H tau, costs): the linear term is shifted so that the model is expressed in
    tau, not in delta-tau, exactly like the reference (:50)."""
    xa, ua = _np(x), _np(u)
    assert xa.shape[0] == ua.shape[0]
    assert xa.shape[1] == ua.shape[1]
    tau = np.concatenate((xa, ua), axis=2)
    grad = _autograd()
    Hs, gs, cs = [], [], []
    for t in range(tau.shape[0]):
        H, g, c0 = _cost_taylor_chainer(Cf, tau[t], grad) if grad else _cost_taylor_fd(Cf, tau[t])
        Hs.append(H)
        gs.append(g - np.einsum("bij,bj->bi", H, tau[t]))
        cs.append(c0)
    return wrap(np.stack(Hs)), wrap(np.stack(gs)), wrap(np.stack(cs))


# -------------------------------------------------------------------------------------------- dynamics
def _fd_column(f0, fp, fm, h):
    """d f / d z_j from f(z), f(z + h e_j), f(z - h e_j): the central difference where the function is smooth.  Where the
    one-sided differences disagree the stencil straddles a kink - in practice a clip inside the dynamics sitting exactly on
    a clamped control - and the one-sided difference of larger magnitude is taken: that is the slope towards the inside of
    the clip, the "inclusive" sub-gradient Chainer's F.clip (and the device pendulum's analytic Jacobian) use."""
    fwd, bwd = (fp - f0) / h, (f0 - fm) / h
    gap = np.abs(fwd - bwd)                                  # smooth: ~h |f''| ~ 1e-5; kink: the jump of the slope
    kink = (gap > 1e-3 * (np.abs(fwd) + np.abs(bwd))) & (gap > 1e-4 * np.maximum(1.0, np.abs(f0)))
    one_sided = np.where(np.abs(fwd) >= np.abs(bwd), fwd, bwd)
    return np.where(kink, one_sided, (fp - fm) / (2.0 * h))


def _jacobian_fd(dynamics, xt, ut):
    """(x_next [B,n], R [B,n,n], S [B,n,m]) of x_next = dynamics(x, u) by central differences."""
    B, n = xt.shape
    m = ut.shape[1]
    nx = _np(dynamics(xt, ut))
    R = np.empty((B, n, n)); S = np.empty((B, n, m))
    hx, hu = _steps(xt, FD_STEP_GRAD), _steps(ut, FD_STEP_GRAD)
    for j in range(n):
        e = np.zeros((B, n)); e[:, j] = hx[:, j]
        R[:, :, j] = _fd_column(nx, _np(dynamics(xt + e, ut)), _np(dynamics(xt - e, ut)), hx[:, j, None])
    for j in range(m):
        e = np.zeros((B, m)); e[:, j] = hu[:, j]
        S[:, :, j] = _fd_column(nx, _np(dynamics(xt, ut + e)), _np(dynamics(xt, ut - e)), hu[:, j, None])
    return nx, R, S


def _jacobian_chainer(dynamics, xt, ut, grad):
    import chainer
    from chainer import functions as F
    xv, uv = chainer.Variable(xt), chainer.Variable(ut)
    nx = dynamics(xv, uv)
    rows = [grad([F.sum(nx[:, j])], [xv, uv]) for j in range(xt.shape[1])]
    return nx.array, np.stack([r[0].array for r in rows], axis=1), np.stack([r[1].array for r in rows], axis=1)


def linearize_dynamics(x, u, dynamics):
    """First-order model x_{t+1} ~ F_t [x_t; u_t] + f_t of the dynamics callable - reference approximate.py:77-119.  Like
    the reference it RE-ROLLS the trajectory from x[0] under u (its :95 `x_ar`) and linearises along that, not along the x
    it was given.  util.LinDx comes back unchanged and the pendulum goes to the device (dmpc_get_traj)."""
    from util import LinDx
    xa, ua = _np(x), _np(u)
    assert xa.shape[0] == ua.shape[0]
    assert xa.shape[1] == ua.shape[1]
    T, B, n = xa.shape
    m = ua.shape[2]
    if isinstance(dynamics, LinDx):
        return dynamics.F, dynamics.f
    from mpc_step import is_pendulum, pendulum_params
    if is_pendulum(dynamics):
        import _native
        ctx = _native.default_context(0)
        Tm = max(T - 1, 1)
        dx = ctx.empty((T, B, n)); Fo = ctx.empty((Tm, B, 3, 4)); fo = ctx.empty((Tm, B, 3))
        ctx.get_traj(np.float64, T, B, n, m, _native.DYN_PENDULUM, ctx.to_device(np.ascontiguousarray(xa[0])),
                     ctx.to_device(np.ascontiguousarray(ua)), None, None, pendulum_params(dynamics), dx, Fo, fo)
        return wrap(Fo.download()[:T - 1]), wrap(fo.download()[:T - 1])
    grad = _autograd()
    xt = xa[0]
    Fs, fs = [], []
    for t in range(T - 1):
        nx, R, S = _jacobian_chainer(dynamics, xt, ua[t], grad) if grad else _jacobian_fd(dynamics, xt, ua[t])
        Fs.append(np.concatenate((R, S), axis=2))
        fs.append(nx - np.einsum("bij,bj->bi", R, xt) - np.einsum("bij,bj->bi", S, ua[t]))
        xt = nx
    if not Fs:
        return wrap(np.zeros((0, B, n, n + m))), wrap(np.zeros((0, B, n)))
    return wrap(np.stack(Fs)), wrap(np.stack(fs))

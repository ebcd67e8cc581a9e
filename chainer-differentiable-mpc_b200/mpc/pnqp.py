"""PNQP on B200 - projected-Newton box QP with the signature of reference mpc/pnqp.py:37.

    x, factor, Index_f, i = PNQP(H, q, lower, upper, x_init=None, n_iter=20)

factor is H_f[B,1,1] when n_dim == 1 and (LU[B,d,d], piv[B,d] int32) otherwise, exactly the
tuple the reference hands to mpc_step (pnqp.py:144).  The LU is computed in fp64 inside the
kernel (csrc/mpc_kernels.cuh, g_pnqp); torch.lu / torch.lu_solve are gone.

coupling: 'batch' reproduces the reference's whole-batch control flow (global convergence test
and shared line-search exit) and needs the batch to fit one CTA; 'element' runs the
reference's algorithm independently per element (== reference with n_batch 1); 'auto' picks
'batch' when it fits, else 'element'.
"""
import os
import sys
import warnings

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.dirname(_here)
for _p in (_pkg,):
    if _p not in sys.path:
        sys.path.append(_p)

import numpy as np  # noqa: E402

import _native  # noqa: E402
from _compat import to_xp, as_f  # noqa: E402

GAMMA = 0.1
DEFAULT_COUPLING = "auto"


def _fits_one_cta(B, m):
    """Whole-batch coupling needs the batch resident in one CTA or in the (up to 16) CTAs of one thread-block cluster."""
    sg = 4 if m <= 4 else (8 if m <= 8 else (16 if m <= 16 else 32))
    return B * sg <= 16 * 1024


def PNQP(H, q, lower, upper, x_init=None, n_iter=20, coupling=None, device=0):
    H, q = as_f(H), as_f(q)
    dt = H.dtype
    lower, upper = as_f(lower, dt), as_f(upper, dt)
    assert (lower <= upper).all(), " lower is larger than upper lower: " + str(lower) + "upper: " + str(upper)
    B, d = H.shape[0], H.shape[1]
    assert list(H.shape) == [B, d, d], "H dim mismatch"
    assert list(q.shape) == [B, d], "q dim mismatch expected" + str([B, d])
    assert list(lower.shape) == [B, d], "lower dim mismatch actual" + str(lower.shape)
    assert list(upper.shape) == [B, d], "upper dim mismatch"
    coupling = coupling or DEFAULT_COUPLING
    auto = coupling == "auto"
    if auto:
        coupling = "batch" if _fits_one_cta(B, d) else "element"
    ctx = _native.default_context(device)
    dev = [ctx.to_device(a) for a in (H, q, lower, upper)]
    dxi = None
    if x_init is not None and to_xp(x_init) is not None:
        dxi = ctx.to_device(as_f(x_init, dt))
    x = ctx.empty((B, d), dt); LU = ctx.empty((B, d, d), dt); piv = ctx.empty((B, d), np.int32)
    free = ctx.empty((B, d), dt); it = ctx.empty((B,), np.int32); fl = ctx.empty((B,), np.int32)
    try:
        ctx.pnqp(dt, B, d, dev[0], dev[1], dev[2], dev[3], dxi, x, LU, piv, free, it, fl, int(n_iter),
                 _native.COUPLING_BATCH if coupling == "batch" else _native.COUPLING_ELEMENT)
    except _native.DiffMpcError as ex:
        if not (auto and coupling == "batch" and "unsupported" in str(ex)):
            raise
        # 'auto' guessed that the batch is resident at once (one CTA / one cluster); the launcher knows better
        ctx.pnqp(dt, B, d, dev[0], dev[1], dev[2], dev[3], dxi, x, LU, piv, free, it, fl, int(n_iter), _native.COUPLING_ELEMENT)
    its = it.download()
    if (fl.download() & _native.FLAG_QP_NOT_CONVERGED).any():
        warnings.warn("Projected Newton Quadratic Programming warning: Did not converge")   # pnqp.py:192
    factor = LU.download() if d == 1 else (LU.download(), piv.download())
    return x.download(), factor, free.download(), int(its.max())

"""Oracle: projected-Newton box QP (test infrastructure).  Follows mpc/pnqp.py:26-201.

minimise 0.5 x^T H x + q^T x  s.t. lower <= x <= upper, batched over B.
Returns (x, factor, index_free, i) with factor = H_f[B,1,1] for n_dim == 1 else
(LU[B,d,d], piv[B,d] int32 1-based) - the same tuple shapes as the reference.
"""
import warnings

import numpy as np

from .linalg import bmv, bger, bquad, bdot, clamp, lu_factor, lu_solve

GAMMA = 0.1        # pnqp.py:23
DECAY = 0.1        # pnqp.py:163
REG = 1e-11        # pnqp.py:73
TOL = 1e-4         # pnqp.py:140
MAX_LS = 10        # pnqp.py:172


def objective(H, q, x):
    return 0.5 * bquad(x, H) + bdot(q, x)      # pnqp.py:26-33


def _pnqp_batch(H, q, lower, upper, x_init, n_iter, lu_fp32):
    B, d = q.shape
    assert (lower <= upper).all(), "lower is larger than upper"
    eye = REG * np.eye(d)[None].repeat(B, axis=0)
    if x_init is None:                                         # pnqp.py:75-83
        if d == 1:
            x0 = -(1.0 / H[:, :, 0]) * q
        else:
            x0 = -lu_solve(lu_factor(H), q, fp32=lu_fp32)
    else:
        x0 = np.array(x_init, copy=True)
    x = clamp(x0, lower, upper)                                # :93
    factor = None
    free = None
    i = 0
    for i in range(n_iter):
        g = bmv(H, x) + q                                      # :98
        act = ((x == lower) & (g > 0.0)) | ((x == upper) & (g < 0.0))   # :110
        free = 1.0 - 1.0 * act
        not_ff = (1.0 - bger(free, free)).astype(bool)
        gf = g.copy()
        gf[act] = 0.0
        Hf = H.copy()
        Hf[not_ff] = 0.0
        Hf += eye                                              # :129
        if d == 1:
            dx = -(1.0 / Hf[:, :, 0]) * gf
            factor = Hf
        else:
            factor = lu_factor(Hf)
            dx = -lu_solve(factor, gf, fp32=lu_fp32)
        large = np.sqrt(np.sum(dx ** 2, axis=1)) >= TOL        # :139-140
        if large.sum() == 0:
            return x, factor, free, i                          # :143-144 (x before dx)
        alpha = np.ones(B, dtype=x.dtype)
        max_lhs = GAMMA
        count = 0
        x_hat = x
        while max_lhs <= GAMMA and count < MAX_LS:             # :172
            x_hat = clamp(x + alpha[:, None] * dx, lower, upper)
            lhs = (GAMMA + 1e-6) * np.ones(B, dtype=x.dtype)
            with np.errstate(divide="ignore", invalid="ignore"):
                ratio = (objective(H, q, x) - objective(H, q, x_hat)) / bdot(g, x - x_hat)
            lhs[large] = ratio[large]
            alpha[lhs <= GAMMA] *= DECAY
            max_lhs = np.max(lhs)
            count += 1
        x = x_hat                                              # :190
    warnings.warn("Projected Newton Quadratic Programming warning: Did not converge")
    return x, factor, free, i


def pnqp(H, q, lower, upper, x_init=None, n_iter=20, lu_fp32=False, coupling="batch"):
    H = np.asarray(H); q = np.asarray(q); lower = np.asarray(lower); upper = np.asarray(upper)
    if coupling == "batch":
        return _pnqp_batch(H, q, lower, upper, x_init, n_iter, lu_fp32)
    assert coupling == "element"
    B, d = q.shape
    xs, frees, its, f0, f1 = [], [], [], [], []
    for b in range(B):
        sl = slice(b, b + 1)
        xi = None if x_init is None else np.asarray(x_init)[sl]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            x, fac, free, i = _pnqp_batch(H[sl], q[sl], lower[sl], upper[sl], xi, n_iter, lu_fp32)
        xs.append(x); frees.append(free); its.append(i)
        if d == 1:
            f0.append(fac)
        else:
            f0.append(fac[0]); f1.append(fac[1])
    factor = np.concatenate(f0) if d == 1 else (np.concatenate(f0), np.concatenate(f1))
    return np.concatenate(xs), factor, np.concatenate(frees), np.array(its, dtype=np.int32)

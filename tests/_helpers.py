"""Shared seeded problem generators for the parity tests (SURVEY.md §8d synthetic inputs)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    d = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    return {k: d[k] for k in d.files}


def rel_err_norm(a, b):
    """Tensor-wide normalised error: max|a-b| / max(1, max|b|).  Kept for quantities whose entries are sums of
    O(1) terms that cancel (finite-difference checks); the parity tests use rel_err below."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b))) / max(1.0, float(np.max(np.abs(b))))


def rel_err(a, b, floor_frac=None):
    """ELEMENT-relative error with an absolute floor: max over entries of |a-b| / (|b| + floor).

    The floor of an entry is floor_frac x the largest magnitude of its own block - one (timestep, batch element)
    block for [T,B,...] tensors whose trailing dims hold at least 8 numbers (gains, C, F, x at n >= 8), otherwise one
    timestep (a [T,B,1] control has no block of its own to be relative to), one row for [B,k >= 8] - and never less than
    1e-6 x the tensor-wide maximum.  A small late-
    horizon gain or a small gradient entry therefore has to be right relative to ITS block, not relative to the largest
    number anywhere in the tensor (the old tensor-wide norm let 1e-6-relative errors in small entries pass).
    floor_frac defaults to 1e-3 for float64 results and 3e-2 for float32 results (entries more than ~30x below their
    block's scale carry no significant digits at 1e-4 in float arithmetic)."""
    a_in = np.asarray(a)
    if floor_frac is None:
        floor_frac = 3e-2 if a_in.dtype == np.float32 else 1e-3
    a = a_in.astype(np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.size == 0:
        return 0.0
    ab = np.abs(b)
    gmax = float(np.max(ab))
    if b.ndim >= 3 and int(np.prod(b.shape[2:])) >= 8:
        sl = np.max(ab, axis=tuple(range(2, b.ndim)), keepdims=True)
    elif b.ndim >= 3:
        sl = np.max(ab, axis=tuple(range(1, b.ndim)), keepdims=True)
    elif b.ndim == 2 and b.shape[1] >= 8:
        sl = np.max(ab, axis=1, keepdims=True)
    else:
        sl = gmax
    floor = np.maximum(np.maximum(floor_frac * sl, 1e-6 * gmax), 1e-300)
    d = np.abs(a - b)
    if not np.all(np.isfinite(d)):
        return float("inf")
    return float(np.max(d / (ab + floor)))


def stable_dynamics(rs, T, B, n, m, rho=0.95, per_t=True):
    """A = I + 0.2 randn rescaled to spectral radius <= rho, B = randn (differentiable_lqr.py:168-171)."""
    A = np.eye(n) + 0.2 * rs.randn(B, n, n)
    for b in range(B):
        r = np.max(np.abs(np.linalg.eigvals(A[b])))
        if r > rho:
            A[b] *= rho / r
    Bm = rs.randn(B, n, m)
    F = np.concatenate((A, Bm), axis=2)
    F = np.repeat(F[None], max(T - 1, 0), axis=0).copy()
    if per_t and T > 1:
        F += 0.01 * rs.randn(*F.shape)
    return F


def psd_cost(rs, T, B, s, sym=True):
    L = rs.randn(T, B, s, s) * 0.3
    C = L @ np.transpose(L, (0, 1, 3, 2)) + np.eye(s)
    if not sym:
        C = C + 0.05 * rs.randn(T, B, s, s)
    c = rs.randn(T, B, s)
    return C, c


def lqr_problem(seed, T, B, n, m, with_f=True, sym=True):
    rs = np.random.RandomState(seed)
    s = n + m
    C, c = psd_cost(rs, T, B, s, sym)
    F = stable_dynamics(rs, T, B, n, m)
    f = 0.1 * rs.randn(max(T - 1, 0), B, n) if with_f else None
    x0 = rs.randn(B, n)
    return dict(x0=x0, C=C, c=c, F=F, f=f, n=n, m=m, T=T, B=B)

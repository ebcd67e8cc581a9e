"""Shared seeded problem generators for the parity tests (SURVEY.md §8d synthetic inputs)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    d = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    return {k: d[k] for k in d.files}


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b))) / max(1.0, float(np.max(np.abs(b))))


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

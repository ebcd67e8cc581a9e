"""Independent analytic pins for the parts of the path the reference's own artefacts leave unpinned.
Linear dynamics + quadratic cost + control box = a convex QP, which is solved here EXACTLY (condensed form, active-set
linear solve, KKT signs checked) without any code shared with the path:

  * forward: the oracle's BoxDDP (box_ddp.py:93-291 -> mpc_step.py:70-328 -> pnqp.py:37-201) must land on that QP's
    solution, with a third of the controls sitting on their bounds, up to the solver's own stopping tolerance
    (PNQP returns before steps shorter than 1e-4, pnqp.py:139-144);
  * backward: the gradients `MPCstep.backward` (mpc_step.py:330-460 -> active_constrained_lqr.py:67-193) returns at the
    solution must equal central finite differences of the imitation loss through the exact QP solution with respect to the
    shared cost parameters (the backward of IL_Env.mpc's repeat of q and p, il_env.py:120-129).

Runs the oracle only (CPU); the CUDA path is compared with the same oracle functions in the -m gpu tests."""
import warnings

import numpy as np

from oracle import boxddp as oddp
from oracle import mpc as ompc

T, B, n, m = 6, 3, 3, 2
s = n + m
LO, HI = -0.5, 0.5


def _problem():
    rs = np.random.RandomState(7)
    L = 0.3 * rs.randn(s, s)
    base = L @ L.T
    A = 0.9 * np.eye(n) + 0.1 * rs.randn(B, n, n)
    F = np.repeat(np.concatenate((A, rs.randn(B, n, m)), axis=2)[None], T - 1, axis=0)
    f = 0.1 * rs.randn(T - 1, B, n)
    x0 = rs.randn(B, n)
    c0 = rs.randn(T, B, s)
    u_exp = np.clip(rs.randn(T, B, m), LO, HI)
    return base, F, f, x0, c0, u_exp


def _cost(q, p, pr):
    C = np.broadcast_to((pr[0] + np.diag(q))[None, None], (T, B, s, s)).copy()
    return C, pr[4] + p[None, None]


def _rollout(F, f, x0, U):
    """tau[T, s] of one element for controls U[T, m]."""
    tau = np.zeros((T, s))
    x = x0
    for t in range(T):
        tau[t, :n], tau[t, n:] = x, U[t]
        if t < T - 1:
            x = F[t] @ tau[t] + f[t]
    return tau


def _exact_qp(C, c, pr, clamp_lo, clamp_hi):
    """Exact minimiser for the given active set; asserts primal feasibility and multiplier signs (KKT)."""
    _, F, f, x0, _, _ = pr
    X, U = np.zeros((T, B, n)), np.zeros((T, B, m))
    for b in range(B):
        d = _rollout(F[:, b], f[:, b], x0[b], np.zeros((T, m))).ravel()
        M = np.zeros((T * s, T * m))
        for j in range(T * m):
            e = np.zeros(T * m); e[j] = 1.0
            M[:, j] = _rollout(F[:, b], f[:, b], x0[b], e.reshape(T, m)).ravel() - d
        Cb = np.zeros((T * s, T * s))
        for t in range(T):
            Cb[t * s:(t + 1) * s, t * s:(t + 1) * s] = 0.5 * (C[t, b] + C[t, b].T)
        H = M.T @ Cb @ M
        h = M.T @ (Cb @ d + c[:, b].ravel())
        lo_set, hi_set = clamp_lo[:, b].ravel(), clamp_hi[:, b].ravel()
        free = ~(lo_set | hi_set)
        u = np.where(lo_set, LO, np.where(hi_set, HI, 0.0))
        u[free] = np.linalg.solve(H[np.ix_(free, free)], -(h[free] + H[np.ix_(free, ~free)] @ u[~free]))
        g = H @ u + h
        assert np.all(u[free] > LO) and np.all(u[free] < HI), "active set: free control outside the box"
        assert np.all(g[lo_set] > 0) and np.all(g[hi_set] < 0), "active set: wrong multiplier sign"
        tau = (M @ u + d).reshape(T, s)
        X[:, b], U[:, b] = tau[:, :n], tau[:, n:]
    return X, U


def _solve_exact(q, p, pr, sets=None):
    C, c = _cost(q, p, pr)
    r = None
    if sets is None:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            r = oddp.box_ddp(pr[3], (C, c), ("linear", pr[1], pr[2]), T, LO, HI, n, m, eps=1e-11, max_iter=60,
                             not_improved_lim=60, ls_decay=0.2, max_ls_iter=10, coupling="element")
        sets = (np.abs(r["u"] - LO) <= 1e-8, np.abs(r["u"] - HI) <= 1e-8)
    X, U = _exact_qp(C, c, pr, *sets)
    return C, c, X, U, sets, r


def _loss(U, pr):
    d = U - pr[5]
    return float(np.mean(d * d))


def test_boxddp_lands_on_the_exact_qp_solution_and_its_gradient_matches_finite_differences():
    pr = _problem()
    q = np.array([1.0, 0.8, 1.2, 0.5, 0.7])
    p = np.array([0.2, -0.1, 0.3, 0.1, -0.2])
    C, c, X, U, sets, r = _solve_exact(q, p, pr)
    clamped = sets[0] | sets[1]
    assert 0.15 < clamped.mean() < 0.7, clamped.mean()              # the bounds are really active
    # forward pin: BoxDDP vs the exact QP solution
    assert np.max(np.abs(r["u"] - U)) < 1e-4 and np.max(np.abs(r["x"] - X)) < 5e-4       # states accumulate the control error
    # backward pin, evaluated at the exact solution
    gu = 2.0 * (U - pr[5]) / U.size
    lo, hi = np.full((T, B, m), LO), np.full((T, B, m), HI)
    dx0, dC, dc, dF, df = ompc.step_backward(C, c, pr[1], pr[2], X, U, lo, hi, None, gu, n, m)
    gq = np.einsum("tbii->i", dC)
    gp = dc.sum(axis=(0, 1))
    h = 1e-5
    for i in range(s):
        e = np.zeros(s); e[i] = h
        fq = (_loss(_solve_exact(q + e, p, pr, sets)[3], pr) - _loss(_solve_exact(q - e, p, pr, sets)[3], pr)) / (2 * h)
        fp = (_loss(_solve_exact(q, p + e, pr, sets)[3], pr) - _loss(_solve_exact(q, p - e, pr, sets)[3], pr)) / (2 * h)
        assert abs(gq[i] - fq) <= 1e-6 * abs(fq) + 1e-9, ("q", i, gq[i], fq)
        assert abs(gp[i] - fp) <= 1e-6 * abs(fp) + 1e-9, ("p", i, gp[i], fp)
    # dF, df, dx0 against finite differences of the same exact solution (one entry each per element)
    def loss_with(Fm=None, fm=None, x0m=None):
        pr2 = (pr[0], pr[1] if Fm is None else Fm, pr[2] if fm is None else fm, pr[3] if x0m is None else x0m, pr[4], pr[5])
        return _loss(_solve_exact(q, p, pr2, sets)[3], pr2)
    for b in range(B):
        for (t, i, j) in ((0, 1, 2), (T - 2, 2, n + 1)):
            E = np.zeros_like(pr[1]); E[t, b, i, j] = h
            fd = (loss_with(Fm=pr[1] + E) - loss_with(Fm=pr[1] - E)) / (2 * h)
            assert abs(dF[t, b, i, j] - fd) <= 1e-6 * abs(fd) + 1e-9, ("F", t, b, i, j, dF[t, b, i, j], fd)
        E = np.zeros_like(pr[2]); E[1, b, 0] = h
        fd = (loss_with(fm=pr[2] + E) - loss_with(fm=pr[2] - E)) / (2 * h)
        assert abs(df[1, b, 0] - fd) <= 1e-6 * abs(fd) + 1e-9, ("f", b, df[1, b, 0], fd)
        E = np.zeros_like(pr[3]); E[b, 2] = h
        fd = (loss_with(x0m=pr[3] + E) - loss_with(x0m=pr[3] - E)) / (2 * h)
        assert abs(dx0[b, 2] - fd) <= 1e-6 * abs(fd) + 1e-9, ("x0", b, dx0[b, 2], fd)

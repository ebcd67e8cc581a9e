"""Oracle: time-varying batched LQR and its KKT adjoint (test infrastructure).

riccati_backward / rollout / lqr_solve  follow lqr/lqr_recursion.py:69-158, :160-200, :202-209
difflqr_backward                        follows lqr/differentiable_lqr.py:78-142
Layout everywhere: C[T,B,s,s] c[T,B,s] F[T-1|T,B,n,s] f[T-1,B,n]|None x0[B,n].
"""
import numpy as np

from .linalg import bmv, bger


def riccati_backward(C, c, F, f, n, m):
    T, B = C.shape[0], C.shape[1]
    Ks = np.empty((T, B, m, n), dtype=C.dtype)
    ks = np.empty((T, B, m), dtype=C.dtype)
    V = v = None
    for t in range(T - 1, -1, -1):
        if t == T - 1:                                   # lqr_recursion.py:81-83
            Q, q = C[t], c[t]
        else:                                            # :85-96
            Ft = F[t]
            FtT = np.transpose(Ft, (0, 2, 1))
            Q = C[t] + (FtT @ V) @ Ft
            if f is None:
                q = c[t] + bmv(FtT, v)
            else:
                q = c[t] + bmv(FtT @ V, f[t]) + bmv(FtT, v)
        Qxx, Qxu = Q[:, :n, :n], Q[:, :n, n:]
        Qux, Quu = Q[:, n:, :n], Q[:, n:, n:]
        qx, qu = q[:, :n], q[:, n:]
        if m == 1:                                       # :112-115 scalar branch
            K = -(1.0 / Quu) * Qux
            k = -(1.0 / Quu[:, :, 0]) * qu
        else:                                            # :116-120 F.batch_inv branch
            Qi = np.linalg.inv(Quu)
            K = -(Qi @ Qux)
            k = -bmv(Qi, qu)
        KT = np.transpose(K, (0, 2, 1))
        Ks[t], ks[t] = K, k
        V = Qxx + Qxu @ K + KT @ Qux + (KT @ Quu) @ K     # :151
        v = qx + bmv(Qxu, k) + bmv(KT, qu) + bmv(KT @ Quu, k)   # :152
    return Ks, ks


def rollout(x0, Ks, ks, F, f, zero_mask=None):
    """lqr_recursion.py:160-200 (zero_mask: active_constrained_lqr.py:172-176)."""
    T = Ks.shape[0]
    B, n = x0.shape
    m = ks.shape[2]
    xs = np.empty((T, B, n), dtype=x0.dtype)
    us = np.empty((T, B, m), dtype=x0.dtype)
    x = x0
    for t in range(T):
        u = bmv(Ks[t], x) + ks[t]
        if zero_mask is not None:
            u = np.where(zero_mask[t], 0.0, u)
        xs[t], us[t] = x, u
        if t < T - 1:
            xn = bmv(F[t], np.concatenate((x, u), axis=1))
            if f is not None:
                xn = xn + f[t]
            x = xn
    return xs, us


def lqr_solve(x0, C, c, F, f, n, m):
    Ks, ks = riccati_backward(C, c, F, f, n, m)
    x, u = rollout(x0, Ks, ks, F, f)
    return x, u, Ks, ks


def difflqr_backward(x0, C, c, F, x, u, gx, gu, n, m, quirk_dC=True, quirk_df=True):
    """differentiable_lqr.py:78-142.  Returns (dx0, dC, dc, dF, df).

    quirk_dC: dC_t = 0.5*(dtau x tau) + (tau x dtau)  (operator precedence, :128);
              False -> the symmetric 0.5*(dtau x tau + tau x dtau).
    quirk_df: df = dlambda[0:T-1] (:133); False -> dlambda[1:T].
    """
    T, B = C.shape[0], C.shape[1]
    s = n + m
    taus = np.concatenate((x, u), axis=2)
    C_Tx = C[T - 1][:, :n, :]
    lam = [None] * T
    lam[T - 1] = bmv(C_Tx, taus[T - 1]) + c[T - 1][:, :n]          # :92
    for t in range(T - 2, -1, -1):                                  # :95-103
        FxT = np.transpose(F[t][:, :n, :n], (0, 2, 1))
        lam[t] = bmv(FxT, lam[t + 1]) + bmv(C[t][:, :n, :], taus[t]) + c[t][:, :n]
    drl = np.concatenate((gx, gu), axis=2)                          # :110
    zf = np.zeros((T - 1, B, n), dtype=C.dtype)
    dx, du, _, _ = lqr_solve(np.zeros_like(x0), C, drl, F, zf, n, m)   # :111-112
    dtaus = np.concatenate((dx, du), axis=2)
    dlam = [None] * T
    dlam[T - 1] = bmv(C_Tx, dtaus[T - 1]) + drl[T - 1][:, :n]       # :115
    for t in range(T - 2, -1, -1):                                  # :117-125
        FxT = np.transpose(F[t][:, :n, :n], (0, 2, 1))
        dlam[t] = bmv(FxT, dlam[t + 1]) + bmv(C[t][:, :n, :], dtaus[t]) + drl[t][:, :n]
    if quirk_dC:
        dC = np.stack([0.5 * bger(dtaus[t], taus[t]) + bger(taus[t], dtaus[t]) for t in range(T)])
    else:
        dC = np.stack([0.5 * (bger(dtaus[t], taus[t]) + bger(taus[t], dtaus[t])) for t in range(T)])
    dc = dtaus.copy()
    if T > 1:
        dF = np.stack([bger(dlam[t + 1], taus[t]) + bger(lam[t + 1], dtaus[t]) for t in range(T - 1)])
        df = np.stack(dlam[:T - 1]) if quirk_df else np.stack(dlam[1:])
    else:
        dF = np.zeros((0, B, n, s), dtype=C.dtype)
        df = np.zeros((0, B, n), dtype=C.dtype)
    return dlam[0], dC, dc, dF, df

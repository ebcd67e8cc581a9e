"""Batched helpers of the oracle (test infrastructure).

Follows util.py:293-298 (xpbmv), :316-329 (xpbger), :349-358 (xpbquad),
:427-434 (xpbdot), :117-123 (xpclamp), :462-482 / :505-528 (torch LU wrappers).
"""
import numpy as np
import torch


def bmv(A, x):
    """[B,p,q] @ [B,q] -> [B,p]"""
    return np.matmul(A, x[:, :, None])[:, :, 0]


def bger(x, y):
    """[B,p] (x) [B,q] -> [B,p,q]"""
    return x[:, :, None] @ y[:, None, :]


def bquad(x, Q):
    """x^T Q x per batch element, evaluated as (x^T Q) x like util.py:356."""
    return ((x[:, None, :] @ Q) @ x[:, :, None])[:, 0, 0]


def bdot(x, y):
    return (x[:, None, :] @ y[:, :, None])[:, 0, 0]


def clamp(x, lo, hi):
    return np.minimum(np.maximum(x, lo), hi)


def lu_factor(A):
    """LAPACK getrf through torch, like util.py:481.  Returns (LU, piv[1-based int32])."""
    LU, piv = torch.linalg.lu_factor(torch.from_numpy(np.ascontiguousarray(A)))
    return LU.numpy(), piv.numpy()


def lu_solve(lu_piv, b, fp32=False):
    """Solve with a factorisation from lu_factor.  fp32=True reproduces the
    reference's float32 cast of LU and rhs (util.py:522-526)."""
    LU, piv = lu_piv
    tb = torch.from_numpy(np.ascontiguousarray(b))
    tLU = torch.from_numpy(np.ascontiguousarray(LU))
    tpiv = torch.from_numpy(np.ascontiguousarray(piv))
    vec = tb.dim() == 2
    if vec:
        tb = tb.unsqueeze(-1)
    if fp32:
        tb = tb.float()
        tLU = tLU.float()
    out = torch.linalg.lu_solve(tLU, tpiv, tb)
    if vec:
        out = out.squeeze(-1)
    return out.numpy()

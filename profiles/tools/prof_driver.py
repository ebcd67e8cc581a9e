#!/usr/bin/env python
"""Minimal (torch-free) driver for ncu: runs the LQR fwd+bwd kernels a few times on a small batch.

    ncu -k regex:'lqr_|adjoint' ... python profiles/tools/prof_driver.py --workload c5 --batch 592 --reps 3
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "chainer-differentiable-mpc_b200"))
import _native  # noqa: E402

SHAPES = {"c5": (32, 8, 100), "c2": (4, 2, 50), "c3": (8, 4, 50), "c4": (3, 1, 20)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--batch", type=int, default=592)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    n, m, T = SHAPES[a.workload]
    B, s = a.batch, n + m
    rs = np.random.RandomState(0)
    ctx = _native.Context(0)
    A = np.eye(n) * 0.9 + 0.02 * rs.randn(n, n)
    Fb = np.concatenate((A, rs.randn(n, m)), axis=1)
    F = np.broadcast_to(Fb, (T - 1, B, n, s)).copy() + 0.001 * rs.randn(1, B, n, s)
    L = 0.3 * rs.randn(s, s)
    C = np.broadcast_to(L @ L.T + np.eye(s), (T, B, s, s)).copy()
    d = dict(x0=rs.randn(B, n), C=C, c=rs.randn(T, B, s), F=F, f=0.1 * rs.randn(T - 1, B, n),
             gx=rs.randn(T, B, n), gu=rs.randn(T, B, m))
    dv = {k: ctx.to_device(v) for k, v in d.items()}
    o = dict(x=ctx.empty((T, B, n)), u=ctx.empty((T, B, m)), Ks=ctx.empty((T, B, m, n)), ks=ctx.empty((T, B, m)),
             fac=ctx.empty((ctx.lqr_fac_elems(T, B, n, m),)), dx0=ctx.empty((B, n)), dC=ctx.empty((T, B, s, s)),
             dc=ctx.empty((T, B, s)), dF=ctx.empty((T - 1, B, n, s)), df=ctx.empty((T - 1, B, n)))
    for _ in range(a.reps):
        ctx.lqr_solve(np.float64, T, B, n, m, dv["x0"], dv["C"], dv["c"], dv["F"], T - 1, dv["f"], o["x"], o["u"],
                      o["Ks"], o["ks"], o["fac"], _native.LQR_FACTOR | _native.LQR_ROLLOUT | _native.LQR_SAVE_FAC)
        ctx.lqr_adjoint(np.float64, T, B, n, m, dv["C"], dv["c"], dv["F"], o["x"], o["u"], dv["gx"], dv["gu"], o["Ks"],
                        o["fac"], o["dx0"], o["dC"], o["dc"], o["dF"], o["df"], _native.ADJ_STRICT_REFERENCE)
    ctx.sync()
    print("ok", float(np.abs(o["x"].download()).max()))


if __name__ == "__main__":
    main()

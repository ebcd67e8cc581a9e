"""LQR_active on B200 - LQR with the controls at active bounds pinned to zero.

Same constructor / methods as reference mpc/active_constrained_lqr.py:16-202; runs
`lqr_solve_kernel` with the MASKED flag (rows/cols of Quu, rows of Qux and entries of qu zeroed,
+1e-8 on the masked diagonal, u_t := 0 where active - reference :112-126, :175).
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.dirname(_here)
for _p in (_pkg,):
    if _p not in sys.path:
        sys.path.append(_p)

import numpy as np  # noqa: E402

import _native  # noqa: E402
from _compat import to_xp, as_f  # noqa: E402


class LQR_active:
    def __init__(self, x_init, C, c, large_f, f, T, n_state, n_ctrl, u_zero_Index=None, device=0):
        assert u_zero_Index is not None, "LQR_active needs u_zero_Index"
        self.T, self.n_state, self.n_ctrl = int(T), int(n_state), int(n_ctrl)
        self.n_sc = self.n_state + self.n_ctrl
        self.x_init = as_f(x_init)
        dt = self.x_init.dtype
        self.C, self.c = as_f(C, dt), as_f(c, dt)
        self.n_batch = self.C.shape[1]
        assert list(self.x_init.shape) == [self.n_batch, self.n_state]
        assert list(self.C.shape) == [self.T, self.n_batch, self.n_sc, self.n_sc], "C dim mismatch"
        assert list(self.c.shape) == [self.T, self.n_batch, self.n_sc], "c dim mismatch"
        self.F = as_f(large_f, dt)
        self.f = None if (f is None or to_xp(f) is None) else as_f(f, dt)
        if self.f is not None:
            assert list(self.f.shape) == [self.T - 1, self.n_batch, self.n_state], " f dim mismatch"
        self.u_zero_Index = np.ascontiguousarray(np.asarray(to_xp(u_zero_Index)).astype(np.uint8))
        assert list(self.u_zero_Index.shape) == [self.T, self.n_batch, self.n_ctrl]
        self._ctx = _native.default_context(device)
        self._res = None

    def _run(self):
        if self._res is None:
            ctx, dt = self._ctx, self.x_init.dtype
            T, B, n, m = self.T, self.n_batch, self.n_state, self.n_ctrl
            d = [ctx.to_device(a) for a in (self.x_init, self.C, self.c, self.F)]
            df = None if self.f is None else ctx.to_device(self.f)
            act = ctx.to_device(self.u_zero_Index)
            x = ctx.empty((T, B, n), dt); u = ctx.empty((T, B, m), dt)
            Ks = ctx.empty((T, B, m, n), dt); ks = ctx.empty((T, B, m), dt)
            ctx.lqr_active_solve(dt, T, B, n, m, d[0], d[1], d[2], d[3], self.F.shape[0], df, act, x, u, Ks, ks)
            self._res = (x.download(), u.download(), Ks.download(), ks.download())
        return self._res

    def backward(self):
        _, _, Ks, ks = self._run()
        return [Ks[t] for t in range(self.T)], [ks[t] for t in range(self.T)]

    def forward(self, Ks=None, ks=None):
        x, u, _, _ = self._run()
        return x, u

    def solve_recursion(self):
        x, u, _, _ = self._run()
        return x, u

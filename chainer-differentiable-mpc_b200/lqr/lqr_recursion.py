"""LqrRecursion on B200 - same constructor and methods as reference lqr/lqr_recursion.py:18-209.

The Riccati sweep (`backward`, reference :69-158) and the rollout (`forward`, :160-200) run in
one CUDA kernel family (csrc/lqr_kernels.cuh, `lqr_solve_kernel`) behind the C ABI entry point
`dmpc_lqr_solve`; `solve_recursion` (:202-209) is a single fused launch.
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
from _compat import to_xp, wrap, as_f  # noqa: E402


class LqrRecursion:
    """Batched time-varying LQR.  x_init[B,n] C[T,B,s,s] c[T,B,s] large_f[T-1|T,B,n,s] f[T-1,B,n]|None."""

    def __init__(self, x_init, C, c, large_f, f, T, n_state, n_ctrl, u_zero_Index=None, device=0):
        if u_zero_Index is not None:
            # reference :121-145 overwrites its own masking and no caller uses it;
            # the working masked solver is active_constrained_lqr.LQR_active.
            raise NotImplementedError("u_zero_Index: use active_constrained_lqr.LQR_active")
        self.T, self.n_state, self.n_ctrl = int(T), int(n_state), int(n_ctrl)
        self.n_sc = self.n_state + self.n_ctrl
        self._x0 = as_f(x_init)
        dt = self._x0.dtype
        self._C = as_f(C, dt)
        self._c = as_f(c, dt)
        self.n_batch = self._C.shape[1]
        assert list(self._x0.shape) == [self.n_batch, self.n_state]
        assert list(self._C.shape) == [self.T, self.n_batch, self.n_sc, self.n_sc], "C dim mismatch"
        assert list(self._c.shape) == [self.T, self.n_batch, self.n_sc], "c dim mismatch"
        F = as_f(large_f, dt) if large_f is not None else np.zeros((0, self.n_batch, self.n_state, self.n_sc), dt)
        assert F.shape[0] in (self.T - 1, self.T) or self.T == 1, "F dimension"
        assert list(F.shape[1:]) == [self.n_batch, self.n_state, self.n_sc], "F dim mismatch"
        self._F = F
        self._f = None
        if f is not None and to_xp(f) is not None:
            self._f = as_f(f, dt)
            assert list(self._f.shape) == [self.T - 1, self.n_batch, self.n_state], " f dim mismatch"
        self._ctx = _native.default_context(device)
        self._dev = None

    # ---- device staging ------------------------------------------------------------------
    def _upload(self):
        if self._dev is None:
            ctx, dt = self._ctx, self._x0.dtype
            T, B, n, m = self.T, self.n_batch, self.n_state, self.n_ctrl
            d = dict(x0=ctx.to_device(self._x0), C=ctx.to_device(self._C), c=ctx.to_device(self._c))
            d["F"] = ctx.to_device(self._F) if self._F.size else ctx.empty((1,), dt)
            d["f"] = ctx.to_device(self._f) if (self._f is not None and self._f.size) else None
            d["Ks"] = ctx.empty((T, B, m, n), dt)
            d["ks"] = ctx.empty((T, B, m), dt)
            d["x"] = ctx.empty((T, B, n), dt)
            d["u"] = ctx.empty((T, B, m), dt)
            self._dev = d
        return self._dev

    def _launch(self, flags):
        d = self._upload()
        F_T = self._F.shape[0] if self.T > 1 else 0
        self._ctx.lqr_solve(self._x0.dtype, self.T, self.n_batch, self.n_state, self.n_ctrl, d["x0"], d["C"], d["c"],
                            d["F"], F_T if F_T else self.T - 1, d["f"], d["x"], d["u"], d["Ks"], d["ks"], None, flags)

    # ---- reference API -------------------------------------------------------------------
    def backward(self):
        """Riccati sweep -> (Ks, ks): lists of T gains [B,m,n] / [B,m] (reference :69-158)."""
        self._launch(_native.LQR_FACTOR)
        d = self._dev
        Ks, ks = d["Ks"].download(), d["ks"].download()
        return [wrap(Ks[t]) for t in range(self.T)], [wrap(ks[t]) for t in range(self.T)]

    def forward(self, Ks, ks):
        """Rollout with the given gains -> (x[T,B,n], u[T,B,m]) (reference :160-200)."""
        assert len(Ks) == self.T, "Ks length error"
        d = self._upload()
        d["Ks"].upload(np.stack([as_f(k, self._x0.dtype) for k in Ks]))
        d["ks"].upload(np.stack([as_f(k, self._x0.dtype) for k in ks]))
        self._launch(_native.LQR_ROLLOUT)
        return wrap(d["x"].download()), wrap(d["u"].download())

    def solve_recursion(self):
        """backward + forward in one launch (reference :202-209)."""
        self._launch(_native.LQR_FACTOR | _native.LQR_ROLLOUT)
        d = self._dev
        return wrap(d["x"].download()), wrap(d["u"].download())

"""Pendulum dynamics marker object for the B200 build (device code: csrc/mpc_kernels.cuh).

Mirrors the solver-facing attributes of reference env_dx/pendulum.py:31-145 (PendulumDx) without
Chainer or matplotlib: params (g, m, l), bounds, mpc_eps, line-search settings and get_true_obj().
`forward(x, u)` evaluates the step on the GPU through dmpc_get_traj.
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


class PendulumDx:
    _dmpc_dynamics = "pendulum"

    def __init__(self, params=None, simple=True):
        assert simple, "only the simple model has device code"
        self.simple = True
        self.max_torque = 2.0
        self.dt = 0.05
        self.n_state = 3
        self.n_ctrl = 1
        self.params = np.array([10.0, 1.0, 1.0]) if params is None else np.asarray(params, dtype=np.float64)
        assert len(self.params) == 3
        self.goal_state = np.array([1.0, 0.0, 0.0])
        self.goal_weights = np.array([1.0, 1.0, 0.1])
        self.ctrl_penalty = 0.001
        self.lower = -2.0
        self.upper = 2.0
        self.mpc_eps = 1e-3
        self.linesearch_decay = 0.2
        self.max_linesearch_iter = 5

    def forward(self, x, u, device=0):
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        squeeze = x.ndim == 1
        if squeeze:
            x, u = x[None], u[None]
        B = x.shape[0]
        ctx = _native.default_context(device)
        out = ctx.empty((2, B, 3))
        uu = np.zeros((2, B, 1)); uu[0] = u
        ctx.get_traj(np.float64, 2, B, 3, 1, _native.DYN_PENDULUM, ctx.to_device(x), ctx.to_device(uu), None, None,
                     tuple(self.params), out)
        r = out.download()[1]
        return r[0] if squeeze else r

    __call__ = forward

    def get_true_obj(self):
        q = np.concatenate((self.goal_weights, self.ctrl_penalty * np.ones(self.n_ctrl)))
        px = -np.sqrt(self.goal_weights) * self.goal_state
        return q, np.concatenate((px, np.zeros(self.n_ctrl)))

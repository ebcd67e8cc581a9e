"""Pendulum dynamics marker object for the B200 build (device code: csrc/mpc_kernels.cuh).

Mirrors the solver-facing attributes of reference env_dx/pendulum.py:31-145 (PendulumDx) without
Chainer or matplotlib: params (g, m, l), bounds, mpc_eps, line-search settings and get_true_obj().
`forward(x, u)` evaluates the step on the GPU through dmpc_get_traj.  The non-`simple` model (damping d, gravity bias b;
pendulum.py:88-93) has no device code: its step is evaluated on the host and the solvers treat the object as an ordinary
callable (plugin path).
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
        self.simple = bool(simple)
        self.max_torque = 2.0
        self.dt = 0.05
        self.n_state = 3
        self.n_ctrl = 1
        default = [10.0, 1.0, 1.0] if self.simple else [10.0, 1.0, 1.0, 0.0, 0.0]      # g, m, l (, d, b)
        self.params = np.array(default) if params is None else np.asarray(getattr(params, "array", params), dtype=np.float64)
        assert len(self.params) == (3 if self.simple else 5)
        self.goal_state = np.array([1.0, 0.0, 0.0])
        self.goal_weights = np.array([1.0, 1.0, 0.1])
        self.ctrl_penalty = 0.001
        self.lower = -2.0
        self.upper = 2.0
        self.mpc_eps = 1e-3
        self.linesearch_decay = 0.2
        self.max_linesearch_iter = 5

    def forward(self, x, u, device=0):
        x = np.ascontiguousarray(getattr(x, "array", x), dtype=np.float64)
        u = np.ascontiguousarray(getattr(u, "array", u), dtype=np.float64)
        squeeze = x.ndim == 1
        if squeeze:
            x, u = x[None], u[None]
        B = x.shape[0]
        if not self.simple:                                   # pendulum.py:88-97, host arithmetic
            g, m, l, d, b = self.params
            uc = np.clip(u, -self.max_torque, self.max_torque)[:, 0]
            th = np.arctan2(x[:, 1], x[:, 0])
            newdth = x[:, 2] + self.dt * (-3.0 * g / (2.0 * l) * (-np.sin(th + b)) + 3.0 * uc / (m * l ** 2) - d * th)
            newth = th + newdth * self.dt
            r = np.stack((np.cos(newth), np.sin(newth), newdth), axis=1)
            return r[0] if squeeze else r
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

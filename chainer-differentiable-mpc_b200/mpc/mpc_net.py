"""MpcNet_dx / MpcNet_cost on B200 - the MPC network links of reference mpc/mpc_net.py:20-227.

    net = MpcNet_dx(T, u_lower, u_upper, n_batch, n_state, n_ctrl, seed, u_init, max_iter=10, ...)
    x, u, costs = net((x_init, QuadCost(C, c)))

A, B are the learned dynamics (same seeding as the reference: `np.random.seed(seed)`, A = I + 0.2 randn, B = randn,
mpc_net.py:57-64); forward broadcasts [A B] over [T-1, B] (util.expand_time_batch) and calls the BoxDDP layer (:72-87).
With real Chainer the gradient reaches A, B through `MPCstep.backward` -> dF -> the backward of F.repeat.  The B200 build
also offers the fused path that never materialises dF[T-1,B,n,s]: `net.backward_numpy(grad_x, grad_u)` runs
`dmpc_mpc_step_backward_reduced` ((T,B)-sum inside the adjoint kernel) and stores dA, dB in `net.A.grad`, `net.B.grad`
- this is also what the forward-only test stub and the Chainer-less numpy mode use.

The reference file defines `MpcNet_dx` twice and a third, identical class under the name `MpcNet_cost` (its body still
learns A, B, mpc_net.py:158-227); both names are provided with that behaviour.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.dirname(_here)
for _p in (_pkg, _here, os.path.join(_pkg, "lqr")):
    if _p not in sys.path:
        sys.path.append(_p)

import numpy as np  # noqa: E402

from _compat import HAVE_CHAINER, NetLinkBase, Parameter, to_xp  # noqa: E402
from box_ddp import BoxDDP  # noqa: E402
from util import expand_time_batch, LinDx  # noqa: E402


class MpcNet_dx(NetLinkBase):
    def __init__(self, T, u_lower, u_upper, n_batch, n_state, n_ctrl, seed, u_init, eps=1e-5, not_improved_lim=5,
                 line_search_decay=0.2, max_line_search_iter=10, best_cost_eps=1e-4, max_iter=10, verbose=False,
                 ilqr_verbose=False, coupling=None, device=0):
        super().__init__()
        self.u_lower, self.u_upper = u_lower, u_upper
        assert (np.asarray(to_xp(u_lower)) <= np.asarray(to_xp(u_upper))).all(), " lower is larger than upper"
        self.T, self.n_batch, self.n_state, self.n_ctrl = T, n_batch, n_state, n_ctrl
        self.n_sc = n_ctrl + n_state
        self.u_init = u_init
        assert list(np.shape(to_xp(u_lower))) == [T, n_batch, n_ctrl], "actual" + str(np.shape(to_xp(u_lower)))
        assert list(np.shape(to_xp(u_upper))) == [T, n_batch, n_ctrl], "actual" + str(np.shape(to_xp(u_upper)))
        with self.init_scope():
            np.random.seed(seed)
            A = np.eye(n_state).astype("float") + 0.2 * np.random.randn(n_state, n_state).astype("float")
            self.A = Parameter(A)
            self.B = Parameter(np.random.randn(n_state, n_ctrl).astype("float"))
        self.mpc_layer = BoxDDP(T=T, u_lower=u_lower, u_upper=u_upper, n_batch=n_batch, n_state=n_state, n_ctrl=n_ctrl,
                                u_init=u_init, eps=eps, not_improved_lim=not_improved_lim,
                                line_search_decay=line_search_decay, max_line_search_iter=max_line_search_iter,
                                best_cost_eps=best_cost_eps, max_iter=max_iter, verbose=verbose,
                                ilqr_verbose=ilqr_verbose, coupling=coupling, device=device)

    def forward(self, inputs):
        x_init, cost = inputs
        if HAVE_CHAINER:
            from chainer import functions as F
            ab_cat = F.concat((self.A, self.B), axis=1)
        else:
            ab_cat = np.concatenate((self.A.array, self.B.array), axis=1)
        large_f = expand_time_batch(ab_cat, self.T - 1, self.n_batch)
        assert list(large_f.shape) == [self.T - 1, self.n_batch, self.n_state, self.n_sc], " Learner's F dimension mismatch"
        f = np.zeros((self.T - 1, self.n_batch, self.n_state), dtype=np.asarray(to_xp(x_init)).dtype)
        return self.mpc_layer((x_init, cost, LinDx(large_f, f)))

    def backward_numpy(self, grad_x, grad_u):
        """Fused-reduction backward of the last forward: returns (dA, dB) and stores them in A.grad / B.grad."""
        g = self.mpc_layer.last_step.backward_reduced_numpy(grad_x, grad_u)     # (dx0, sum dC, sum dc, sum dF, sum df)
        dF = g[3]
        self.A.grad, self.B.grad = dF[:, :self.n_state].copy(), dF[:, self.n_state:].copy()
        return self.A.grad, self.B.grad


class MpcNet_cost(MpcNet_dx):
    """Same body as MpcNet_dx in the reference (mpc_net.py:158-227)."""

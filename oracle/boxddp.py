"""Oracle: box-DDP outer loop (test infrastructure).  Follows mpc/box_ddp.py:93-291.

cost = (C, c) QuadCost; dynamics = ('linear', F, f) or ('pendulum', params).
Returns dict(x, u, costs, n_iter, status, full_du_norm_best, full_du_norm_last,
detach_mask, F_lin, f_lin) - F_lin/f_lin are the linearisation at the returned
point, i.e. what the final no-op MPCstep retains for the adjoint (:235-259).
"""
import numpy as np

from . import mpc as _mpc
from . import pendulum as _pend


def _linearise(x, u, dynamics):
    if dynamics[0] == "linear":
        return dynamics[1], dynamics[2]
    return _pend.linearize(x[0], u, dynamics[1])


def box_ddp(x_init, cost, dynamics, T, u_lower, u_upper, n, m, u_init=None, eps=1e-5,
            not_improved_lim=5, ls_decay=0.2, max_ls_iter=10, best_cost_eps=1e-4, max_iter=10,
            lu_fp32=False, coupling="batch"):
    B = x_init.shape[0]
    C, c = cost
    if np.isscalar(u_lower):                                        # :68-90 (Q9)
        u_lower = np.full((T, B, m), float(u_lower))
        u_upper = np.full((T, B, m), float(u_upper))
    if u_init is None:
        u = np.zeros((T, B, m), dtype=x_init.dtype)
    elif list(u_init.shape) == [T, m]:
        u = np.repeat(u_init[:, None, :], B, axis=1)
    else:
        u = np.array(u_init, copy=True)
    best = None
    n_not_improved = 0
    status = "max_iter"
    fo = None
    it = 0
    for it in range(max_iter):
        x = _mpc.get_traj(x_init, u, dynamics)                      # :123
        F, f = _linearise(x, u, dynamics)                           # :127-131
        x, u, fo, _ = _mpc.step_forward(C, c, F, f, x, u, u_lower, u_upper, cost, dynamics,
                                        ls_decay, max_ls_iter, n, m, need_expand=True,
                                        lu_fp32=lu_fp32, coupling=coupling)     # :160-172
        n_not_improved += 1
        if best is None:
            best = dict(x=x.copy(), u=u.copy(), costs=fo.costs.copy(), full_du_norm=fo.full_du_norm.copy())
        else:
            for j in range(B):                                      # :200-209
                if fo.costs[j] <= best["costs"][j] + best_cost_eps:
                    n_not_improved = 0
                    best["x"][:, j] = x[:, j]
                    best["u"][:, j] = u[:, j]
                    best["costs"][j] = fo.costs[j]
                    best["full_du_norm"][j] = fo.full_du_norm[j]
        if max(fo.full_du_norm) < eps:                              # :223
            status = "converged"
            break
        if n_not_improved > not_improved_lim:                       # :226
            status = "not_improved"
            break
    x, u = best["x"], best["u"]
    F, f = _linearise(x, u, dynamics)                               # :235-238
    detach = None
    if max(best["full_du_norm"]) > eps:                             # :265-274
        detach = fo.full_du_norm < eps
    return dict(x=x, u=u, costs=best["costs"], n_iter=it + 1, status=status,
                full_du_norm_best=best["full_du_norm"], full_du_norm_last=fo.full_du_norm,
                detach_mask=detach, F_lin=F, f_lin=f, u_lower=u_lower, u_upper=u_upper)

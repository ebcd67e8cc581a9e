"""BoxDDP on B200 - the box-constrained iLQR outer loop with the API of reference mpc/box_ddp.py:24-291.

    solver = BoxDDP(T, u_lower, u_upper, n_batch, n_state, n_ctrl, u_init, eps=..., max_iter=..., ...)
    x, u, costs = solver((x_init, QuadCost(C, c), dynamics))

Per iteration (reference :121-230): rollout of the nominal controls (`dmpc_get_traj`), linearisation
(LinDx: as given; pendulum: analytic Jacobian on the device, replacing approximate.linearize_dynamics; any other
callable: approximate.linearize_dynamics / approximate_cost on the host, then MPCstep's plugin path), one fused MPC step (`dmpc_mpc_step_forward`), then the reference's bookkeeping: per-element best
trajectory (vectorised instead of the Python loop over B at :200-209, same semantics), global exits on
max(full_du_norm) < eps and on the shared n_not_improved counter, and the final no-op MPCstep whose
backward is the gradient path (:247-259).  The batch-coupled quirks (H2 iii, iv) are kept on the host.
"""
import copy
import os
import sys
import warnings

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.dirname(_here)
for _p in (_pkg, _here, os.path.join(_pkg, "lqr")):
    if _p not in sys.path:
        sys.path.append(_p)

import numpy as np  # noqa: E402

import _native  # noqa: E402
from _compat import LinkBase, to_xp, wrap, as_f  # noqa: E402
from mpc_step import MPCstep, is_pendulum, pendulum_params  # noqa: E402
from util import QuadCost, LinDx  # noqa: E402


class BoxDDP(LinkBase):
    def __init__(self, T, u_lower, u_upper, n_batch, n_state, n_ctrl, u_init, eps=1e-5, not_improved_lim=5,
                 line_search_decay=0.2, max_line_search_iter=10, best_cost_eps=1e-4, max_iter=10,
                 detach_unconverged=True, exit_unconverged=True, verbose=False, ilqr_verbose=False,
                 update_dynamics=True, coupling=None, device=0, device_loop=True):
        if LinkBase is not object:
            super().__init__()
        self.T, self.n_batch, self.n_state, self.n_ctrl = int(T), int(n_batch), int(n_state), int(n_ctrl)
        self.n_sc = self.n_state + self.n_ctrl
        self.eps = eps
        self.not_improved_lim = not_improved_lim
        self.ls_decay = line_search_decay
        self.max_ls_iter = max_line_search_iter
        self.best_cost_eps = best_cost_eps
        self.max_iter = max_iter
        self.verbose = verbose
        self.ilqr_verbose = ilqr_verbose
        self.u_init = u_init
        self.detach_unconverged = detach_unconverged
        self.exit_unconverged = exit_unconverged
        self.update_dynamics = update_dynamics
        self.coupling = coupling
        self.device = device
        # device_loop: run the whole iLQR loop through dmpc_boxddp_solve (tensors stay in HBM, one 32-byte status
        # read per iteration).  verbose=True needs the per-iteration table of the reference and uses the host loop.
        self.device_loop = device_loop
        shape = (self.T, self.n_batch, self.n_ctrl)
        if isinstance(u_lower, float) or np.isscalar(u_lower):      # reference :68-90 (Q9)
            self.u_lower = np.full(shape, float(u_lower))
            self.u_upper = np.full(shape, float(u_upper))
        else:
            self.u_lower = np.asarray(to_xp(u_lower), dtype=np.float64)
            self.u_upper = np.asarray(to_xp(u_upper), dtype=np.float64)
            assert list(self.u_lower.shape) == list(shape), "actual" + str(self.u_lower.shape)
            assert list(self.u_upper.shape) == list(shape)
        self.last_step = None
        self.info = None
        self._dev_cache = None
        self.u_device = None

    # ---- helpers -----------------------------------------------------------------------------
    def _rollout(self, ctx, x_init, u, dynamics):
        """x = get_traj(u) and (F, f) at (x, u) -- reference :123-131."""
        T, B, n, m = self.T, self.n_batch, self.n_state, self.n_ctrl
        dt = np.float64
        dx = ctx.empty((T, B, n), dt)
        if isinstance(dynamics, LinDx):
            F, f = as_f(dynamics.F, dt), (None if to_xp(dynamics.f) is None else as_f(dynamics.f, dt))
            ctx.get_traj(dt, T, B, n, m, _native.DYN_LINEAR, ctx.to_device(x_init), ctx.to_device(u), ctx.to_device(F),
                         None if f is None else ctx.to_device(f), None, dx)
            return dx.download(), F, f
        if is_pendulum(dynamics):
            Fo = ctx.empty((max(T - 1, 1), B, 3, 4), dt); fo = ctx.empty((max(T - 1, 1), B, 3), dt)
            ctx.get_traj(dt, T, B, n, m, _native.DYN_PENDULUM, ctx.to_device(x_init), ctx.to_device(u), None, None,
                         pendulum_params(dynamics), dx, Fo, fo)
            return dx.download(), Fo.download()[:T - 1], fo.download()[:T - 1]
        # plugin: roll out and linearise through the Python callable on the host (reference :123, :129)
        from approximate import linearize_dynamics
        from util import xpget_traj
        x = xpget_traj(T, u, x_init, dynamics)
        F, f = linearize_dynamics(x, u, dynamics)
        return np.asarray(x, dtype=dt), as_f(F, dt), as_f(f, dt)

    def _solve_on_device(self, ctx, x_init, C_arr, c_arr, true_dyn, u):
        """The loop of reference :121-230 in one C-ABI call.  Returns (best, du_last, n_iter, status, F_lin, f_lin)."""
        from mpc_step import resolve_coupling, MAX_LS_TRIALS
        T, B, n, m, s = self.T, self.n_batch, self.n_state, self.n_ctrl, self.n_sc
        dt = np.float64
        Tm = max(T - 1, 1)
        if isinstance(true_dyn, LinDx):
            dyn, params = _native.DYN_LINEAR, None
            F_h, f_h = true_dyn.F, true_dyn.f
            dF, F_T = ctx.to_device(F_h), F_h.shape[0]
            df = None if f_h is None else ctx.to_device(f_h[:T - 1])
            F_lin = f_lin = None
        elif is_pendulum(true_dyn):
            dyn, params = _native.DYN_PENDULUM, pendulum_params(true_dyn)
            dF = df = None
            F_T = T - 1
            F_lin, f_lin = ctx.empty((Tm, B, n, s), dt), ctx.empty((Tm, B, n), dt)
        else:
            raise NotImplementedError("dynamics must be util.LinDx or a PendulumDx (SURVEY.md H3)")
        coupling = resolve_coupling(self.coupling, B, n, m)
        o = dict(x=ctx.empty((T, B, n), dt), u=ctx.empty((T, B, m), dt), costs=ctx.empty((B,), dt),
                 du=ctx.empty((B,), dt), du_last=ctx.empty((B,), dt))
        dev_in = [a if isinstance(a, _native.DeviceArray) else ctx.to_device(a)
                  for a in (x_init, C_arr, c_arr, self.u_lower, self.u_upper, u)]

        def solve(cpl):
            return ctx.boxddp_solve(
                dt, T, B, n, m, dev_in[0], dev_in[1], dev_in[2], dev_in[3], dev_in[4], dyn, dF, F_T, df, params, dev_in[5],
                self.eps, self.best_cost_eps, self.ls_decay, self.not_improved_lim, self.max_iter, MAX_LS_TRIALS,
                _native.COUPLING_BATCH if cpl == "batch" else _native.COUPLING_ELEMENT,
                o["x"], o["u"], o["costs"], o["du"], o["du_last"], F_lin, f_lin)
        try:
            n_iter, status, flags = solve(coupling)
        except _native.DiffMpcError as ex:      # 'auto' guessed the batch is resident at once; the launcher knows better
            from mpc_step import DEFAULT_COUPLING
            if not ((self.coupling or DEFAULT_COUPLING) == "auto" and coupling == "batch" and "unsupported" in str(ex)):
                raise
            n_iter, status, flags = solve("element")
        if flags & _native.FLAG_QP_NOT_CONVERGED:
            warnings.warn("Projected Newton Quadratic Programming warning: Did not converge")
        if flags & _native.FLAG_LS_CAPPED:
            warnings.warn("MPCstep line search hit the %d-trial cap" % MAX_LS_TRIALS)
        best = dict(x=o["x"].download(), u=o["u"].download(), costs=o["costs"].download(), full_du_norm=o["du"].download())
        self.u_device = o["u"]          # best controls, still in HBM: what _native.WarmStartCache.put takes
        # the final no-op MPCstep differentiates at exactly these tensors: keep the device copies for its backward
        self._dev_cache = dict(C=dev_in[1], c=dev_in[2], lo=dev_in[3], hi=dev_in[4], x=o["x"], u=o["u"],
                               F=(F_lin if F_lin is not None else dF))
        if F_lin is not None:
            large_f, f = F_lin.download()[:T - 1], f_lin.download()[:T - 1]
        else:
            large_f, f = true_dyn.F, true_dyn.f
        return best, o["du_last"].download(), n_iter, {0: "max_iter", 1: "converged", 2: "not_improved"}[status], large_f, f

    def forward(self, inputs):
        x_init, cost, dynamics = inputs
        T, B, n, m = self.T, self.n_batch, self.n_state, self.n_ctrl
        x_init = as_f(x_init, np.float64)
        assert list(x_init.shape) == [B, n], " x_init dim mismatch"
        quad = isinstance(cost, QuadCost)
        if not quad and not callable(cost):
            raise TypeError("cost must be a util.QuadCost or a callable tau[B,s] -> cost[B]")
        plugin = not quad or not (isinstance(dynamics, LinDx) or is_pendulum(dynamics))
        if plugin and not (isinstance(dynamics, LinDx) or callable(dynamics)):
            raise TypeError("dynamics must be a util.LinDx, a PendulumDx or a callable (x[B,n], u[B,m]) -> x_next[B,n]")
        ctx = _native.default_context(self.device)
        u_dev = None
        if isinstance(self.u_init, _native.DeviceArray):            # warm start already in HBM (_native.WarmStartCache)
            u_dev = self.u_init
            assert list(u_dev.shape) == [T, B, m] and u_dev.dtype == np.float64, "device u_init must be float64 [T,B,m]"
            u = None
        elif self.u_init is None:
            u = np.zeros((T, B, m))
        else:
            u = np.asarray(to_xp(self.u_init), dtype=np.float64)
            if list(u.shape) == [T, m]:
                u = np.repeat(u[:, None, :], B, axis=1)
        if u is not None:
            assert list(u.shape) == [T, B, m], "u dim mismatch, actual" + str(u.shape)
            assert not np.isnan(u).any()
        if quad:
            C_arr, c_arr = as_f(cost.C, np.float64), as_f(cost.c, np.float64)
            true_cost = QuadCost(C_arr, c_arr)
        else:
            from approximate import approximate_cost
            true_cost = cost
        if isinstance(dynamics, LinDx):
            true_dyn = LinDx(as_f(dynamics.F, np.float64), None if to_xp(dynamics.f) is None else as_f(dynamics.f, np.float64))
        else:
            true_dyn = dynamics
        best = None
        n_not_improved = 0
        for_out = None
        status = "max_iter"
        n_iter = 0
        on_device = self.device_loop and not self.verbose and not self.ilqr_verbose and not plugin
        if u is None and not on_device:
            u = u_dev.download()                                       # the host loop iterates on host arrays
        if on_device:
            best, du_last, n_iter, status, large_f, f = self._solve_on_device(ctx, x_init, C_arr, c_arr, true_dyn,
                                                                             u_dev if u_dev is not None else u)
            print({"converged": "Converged", "not_improved": "Not improved lim", "max_iter": "Not Converged "}[status])
        for i in range(0 if on_device else self.max_iter):
            n_iter = i + 1
            x, large_f, f = self._rollout(ctx, x_init, u, true_dyn)
            if not quad:                                                            # reference :133-136
                C_arr, c_arr = (as_f(v, np.float64) for v in approximate_cost(x, u, cost)[:2])
            step = MPCstep(controls=u, T=T, u_upper=self.u_upper, u_lower=self.u_lower, n_batch=B, n_state=n,
                           n_ctrl=m, current_states=x, true_cost=true_cost, true_dynamics=true_dyn,
                           ls_decay=self.ls_decay, max_ls_iter=self.max_ls_iter, verbose=self.ilqr_verbose,
                           need_expand=True, coupling=self.coupling, device=self.device)
            x, u = step._forward_arrays(C_arr, c_arr, large_f, f)
            back_out, for_out = step.back_out, step.for_out
            n_not_improved += 1
            if best is None:
                best = dict(x=x.copy(), u=u.copy(), costs=for_out.costs.copy(), full_du_norm=for_out.full_du_norm.copy())
            else:
                better = for_out.costs <= best["costs"] + self.best_cost_eps      # reference :200-209
                if better.any():
                    n_not_improved = 0
                    best["x"][:, better] = x[:, better]
                    best["u"][:, better] = u[:, better]
                    best["costs"][better] = for_out.costs[better]
                    best["full_du_norm"][better] = for_out.full_du_norm[better]
            if self.verbose:
                print("| iter %d | mean(cost) %.4e | ||full_du||_max %.2e | mean(alphas) %.2e | total_qp_iters %d |" % (
                    i, np.mean(best["costs"]), np.max(for_out.full_du_norm), for_out.mean_alphas, back_out.n_total_qp_iter))
            if max(for_out.full_du_norm) < self.eps:                               # reference :223-230
                print("Converged")
                status = "converged"
                break
            if n_not_improved > self.not_improved_lim:
                print("Not improved lim")
                status = "not_improved"
                break
            if i == self.max_iter - 1:
                print("Not Converged ")
        x, u = best["x"], best["u"]
        # linearise at the returned point (reference :235-242) and attach the differentiable graph
        if not on_device:
            _, large_f, f = self._rollout(ctx, x[0], u, true_dyn)
            du_last = for_out.full_du_norm
            if not quad:                                                            # reference :241-242
                C_arr, c_arr = (as_f(v, np.float64) for v in approximate_cost(x, u, cost)[:2])
        final = MPCstep(controls=u, T=T, u_upper=self.u_upper, u_lower=self.u_lower, n_batch=B, n_state=n, n_ctrl=m,
                        current_states=x, true_cost=true_cost, true_dynamics=true_dyn, ls_decay=self.ls_decay,
                        max_ls_iter=self.max_ls_iter, verbose=self.ilqr_verbose, need_expand=True,
                        no_op_forward=True, device=self.device)
        if isinstance(dynamics, LinDx):
            F_in, f_in = dynamics.F, dynamics.f
        else:
            F_in, f_in = large_f, f
        if self.update_dynamics:                                                   # reference :252-258
            C_in, c_in = C_arr, c_arr
        else:
            C_in, c_in = (cost.C, cost.c) if quad else (wrap(C_arr), wrap(c_arr))
            F_in, f_in = to_xp(F_in), to_xp(f_in)
        if on_device and quad:
            final._dev_cache = self._dev_cache       # same values as the host arrays handed to apply() below
        out = final.apply((x[0].copy(), C_in, c_in, F_in, f_in))
        x_new, u_new = out[0], out[1]
        self.last_step = final
        detach_mask = None
        if self.detach_unconverged and max(best["full_du_norm"]) > self.eps:       # reference :263-289
            if self.verbose:
                print("LQR Warning: All examples did not converge to a fixed point.")
                print("Detaching and *not* backpropping through the bad examples.")
            warnings.warn("LQR Warning: All examples did not converge to a fixed point.")
            detach_mask = du_last < self.eps
            Ix = np.broadcast_to(detach_mask[None, :, None], (T, B, n)).astype(np.float64)
            Iu = np.broadcast_to(detach_mask[None, :, None], (T, B, m)).astype(np.float64)
            x_new = x_new * Ix + copy.deepcopy(to_xp(x_new)) * (1.0 - Ix)
            u_new = u_new * Iu + copy.deepcopy(to_xp(u_new)) * (1.0 - Iu)
        self.info = dict(n_iter=n_iter, status=status, detach_mask=detach_mask, full_du_norm_best=best["full_du_norm"],
                         full_du_norm_last=du_last, F_lin=large_f, f_lin=f)
        return x_new, u_new, best["costs"]

    if LinkBase is object:
        def __call__(self, inputs):
            return self.forward(inputs)

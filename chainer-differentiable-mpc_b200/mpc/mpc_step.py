"""MPCstep on B200 - one box-DDP / iLQR step as a FunctionNode, API of reference mpc/mpc_step.py:33-460.

forward  (reference :288-328): Taylor shift + bounded Riccati sweep with one PNQP per timestep +
          line search through the true dynamics/cost - ONE kernel launch (`mpc_forward_kernel`).
backward (reference :330-460): active-set LQR (LQR_active) + lambda/d-lambda recursions + outer
          products - three launches (`dmpc_mpc_step_backward`).

Fused path: true_cost = util.QuadCost and true_dynamics = util.LinDx or a pendulum object (the reference's
env_dx.pendulum.PendulumDx or pendulum_dx.PendulumDx of this package), whose step and analytic Jacobian are device
code.  Plugin path: any other Python callable (cost: tau [B,s] -> [B]; dynamics: (x [B,n], u [B,m]) -> [B,n]) cannot
run inside a kernel, so backward_rec still runs on the GPU (the same kernel, max_ls_trials < 0 = "sweep only") and the
line search of forward_rec (:175-286) runs on the host through the callable, as the reference does.

The batch-scrambled `full_du_norm` / `alpha_du_norm` (reference :261-263, :275-277) are reproduced
bit-faithfully on the host from the kernel's alpha=1 controls.
"""
import os
import sys
import warnings
from collections import namedtuple

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.dirname(_here)
for _p in (_pkg, _here, os.path.join(_pkg, "lqr")):
    if _p not in sys.path:
        sys.path.append(_p)

import numpy as np  # noqa: E402

import _native  # noqa: E402
from _compat import FunctionNodeBase, to_xp, wrap, as_f  # noqa: E402
from util import QuadCost, LinDx  # noqa: E402

LqrBackOut = namedtuple("lqrBackOut", "n_total_qp_iter")
_LqrForOutT = namedtuple("lqrForOut", "objs full_du_norm alpha_du_norm mean_alphas costs")


class LqrForOut:
    """The reference's lqrForOut namedtuple (mpc_step.py:25-30), with the two batch-scrambled step norms evaluated on first
    use: they are two transposed reductions over [T,B,m] that a caller of a single MPC step (and the latency path) often
    never reads.  Attribute access, indexing, iteration and len() behave like the namedtuple's."""
    _fields = _LqrForOutT._fields

    def __init__(self, objs, full_du_norm, alpha_du_norm, mean_alphas, costs):
        self.objs, self.mean_alphas, self.costs = objs, mean_alphas, costs
        self._full, self._alpha = full_du_norm, alpha_du_norm          # arrays, or zero-argument callables

    @property
    def full_du_norm(self):
        if callable(self._full):
            self._full = self._full()
        return self._full

    @property
    def alpha_du_norm(self):
        if callable(self._alpha):
            self._alpha = self._alpha()
        return self._alpha

    def _astuple(self):
        return _LqrForOutT(self.objs, self.full_du_norm, self.alpha_du_norm, self.mean_alphas, self.costs)

    def __iter__(self):
        return iter(self._astuple())

    def __getitem__(self, i):
        return self._astuple()[i]

    def __len__(self):
        return 5

    def __repr__(self):
        return repr(self._astuple())

DEFAULT_COUPLING = "auto"     # 'batch' | 'element' | 'auto' (see pnqp.py)
MAX_LS_TRIALS = 64            # safety cap of the per-element line search (reference has none, Q5)


def is_pendulum(dyn):
    """The package's pendulum_dx.PendulumDx (explicit `_dmpc_dynamics` marker) or the reference's
    env_dx.pendulum.PendulumDx (a chainer.Link: recognised by its class name AND the attributes the device code is a
    restatement of - n_state 3, n_ctrl 1, params, dt, max_torque); an unrelated class that merely shares the name is
    not accepted, and the non-`simple` model (damping / gravity bias, pendulum.py:88-93) has no device code: it is an
    ordinary callable and takes the plugin path."""
    if hasattr(dyn, "simple") and not dyn.simple:
        return False
    if getattr(dyn, "_dmpc_dynamics", None) == "pendulum":
        return True
    return (type(dyn).__name__ == "PendulumDx" and getattr(dyn, "n_state", None) == 3 and getattr(dyn, "n_ctrl", None) == 1
            and all(hasattr(dyn, a) for a in ("params", "dt", "max_torque")))


def pendulum_params(dyn):
    """(g, m, l, dt, max_torque) handed to the device step (env_dx/pendulum.py:40-41,81-97)."""
    p = np.asarray(to_xp(dyn.params), dtype=np.float64).ravel()
    dt, maxu = float(getattr(dyn, "dt", 0.05)), float(getattr(dyn, "max_torque", 2.0))
    assert dt > 0 and maxu > 0
    return (float(p[0]), float(p[1]), float(p[2]), dt, maxu)


def _group_size(n, m):
    table = {(3, 1): 4, (4, 2): 8, (8, 4): 16}
    if (n, m) in table:
        return table[(n, m)]
    s = n + m
    return 8 if (s <= 6 and m <= 8) else (16 if (s <= 14 and m <= 16) else (32 if s <= 24 else 256))


def resolve_coupling(coupling, B, n, m):
    """'auto' = the reference's literal whole-batch control flow whenever the batch can be resident at once - in one CTA
    or spread over the (up to 16) CTAs of one thread-block cluster (csrc/mpc_launch.cu launch_elems / launch_mpc_tpe) -
    and the per-element control flow (== reference with n_batch 1) beyond that."""
    coupling = coupling or DEFAULT_COUPLING
    if coupling == "auto":
        if m == 1 and n in (2, 3):                       # thread-per-element kernel: 256 elements per CTA
            return "batch" if B <= 16 * 256 else "element"
        G = _group_size(n, m)
        if G > 32:
            return "element"
        per_elem = (n + m) ** 2 * 8 * 6                  # shared-memory bytes per element, roughly (mpc_layout)
        epb = max(1, min(1024 // G, (200 * 1024) // per_elem))
        return "batch" if B <= 16 * epb else "element"
    return coupling


def scrambled_norm(du, B, T, m):
    """reference :261-263: transpose(0,2,1).reshape(B, T*m) mixes batch elements (kept as is)."""
    d = np.transpose(du, (0, 2, 1)).reshape(B, T * m)
    return np.sqrt(np.sum(d ** 2, axis=1))


class MPCstep(FunctionNodeBase):
    def __init__(self, controls, T, u_upper, u_lower, n_batch, n_state, n_ctrl, current_states,
                 true_cost, true_dynamics, ls_decay, max_ls_iter, verbose=False, need_expand=False,
                 no_op_forward=False, coupling=None, device=0):
        super().__init__()
        self.controls = to_xp(controls)
        self.u_upper = to_xp(u_upper)
        self.u_lower = to_xp(u_lower)
        self.n_state, self.n_ctrl, self.n_batch, self.T = int(n_state), int(n_ctrl), int(n_batch), int(T)
        self.n_sc = self.n_state + self.n_ctrl
        self.verbose = verbose
        self.back_out = None
        self.for_out = None
        self.current_states = to_xp(current_states)
        self.true_cost = true_cost
        self.true_dynamics = true_dynamics
        self.need_expand = need_expand
        self.ls_decay = ls_decay
        self.max_ls_iter = max_ls_iter
        self.no_op_forward = no_op_forward
        self.coupling = coupling
        self._ctx = _native.default_context(device)
        self._fwd = None       # retained host copies for backward
        self._dev_cache = None # device copies of (C, c, F, x, u, lo, hi) a caller already holds (BoxDDP's device loop)
        self.aux = None        # extra kernel outputs (Ks, ks, alphas, free masks, ...)

    # ---- forward ---------------------------------------------------------------------------
    def _forward_arrays(self, C_hat, c_hat, F_hat, f_hat):
        T, B, n, m, s = self.T, self.n_batch, self.n_state, self.n_ctrl, self.n_sc
        ctx = self._ctx
        C_hat = as_f(C_hat)
        dt = C_hat.dtype
        c_hat, F_hat = as_f(c_hat, dt), as_f(F_hat, dt)
        assert list(C_hat.shape) == [T, B, s, s], "C hat dim mismatch"
        assert list(c_hat.shape) == [T, B, s], str(c_hat.shape) + " c hat dim mismatch: expected " + str([T, B, s])
        assert F_hat.shape[0] in (T, T - 1), "F_hat dimension"
        assert list(F_hat.shape[1:]) == [B, n, s], str(F_hat.shape) + "F_hat dim mismatch"
        if f_hat is not None:
            f_hat = as_f(f_hat, dt)
            assert list(f_hat.shape) in ([T - 1, B, n], [T, B, n]), " f_hat dim mismatch"
        u_nom, x_nom = as_f(self.controls, dt), as_f(self.current_states, dt)
        lo, hi = as_f(self.u_lower, dt), as_f(self.u_upper, dt)
        assert not (np.isnan(np.min(u_nom)) or np.isnan(np.min(lo)) or np.isnan(np.min(hi)))    # min propagates NaN
        assert (lo <= hi).all(), " lower is larger than upper"
        if self._is_plugin():
            return self._forward_plugin(C_hat, c_hat, F_hat, f_hat, x_nom, u_nom, lo, hi)
        tC, tc = as_f(self.true_cost.C, dt), as_f(self.true_cost.c, dt)
        # host -> device: every input that is not an alias of another one; small problems go through ONE packed copy
        ins = {"C": C_hat, "c": c_hat, "F": F_hat, "x_nom": x_nom, "u_nom": u_nom, "lo": lo, "hi": hi}
        if not (f_hat is None or self.need_expand):
            ins["f"] = f_hat[:T - 1]
        if not (tC is C_hat or np.shares_memory(tC, C_hat)):
            ins["tC"] = tC
        if not (tc is c_hat or np.shares_memory(tc, c_hat)):
            ins["tc"] = tc
        if isinstance(self.true_dynamics, LinDx):
            dyn, params = _native.DYN_LINEAR, None
            tF = as_f(self.true_dynamics.F, dt)
            tf = None if to_xp(self.true_dynamics.f) is None else as_f(self.true_dynamics.f, dt)
            if not np.shares_memory(tF, F_hat):
                ins["tF"] = tF
            if tf is not None:
                ins["tf"] = tf
        else:
            dyn, params = _native.DYN_PENDULUM, pendulum_params(self.true_dynamics)
            tf = None
        def out_specs():
            return [("x", (T, B, n), dt), ("u", (T, B, m), dt), ("Ks", (T, B, m, n), dt), ("ks", (T, B, m), dt),
                    ("u_first", (T, B, m), dt), ("objs", (T, B), dt), ("costs", (B,), dt), ("old", (B,), dt),
                    ("alphas", (B,), dt), ("n_qp", (T, B), np.int32), ("free", (T, B, m), np.uint8),
                    ("n_ls", (B,), np.int32), ("flags", (B,), np.int32)]
        packed = sum(a.nbytes for a in ins.values()) <= _native.PACK_LIMIT_BYTES
        if packed:
            lkey = (T, B, n, m, dt.char, F_hat.shape[0])
            pin = _native.PackedBuffers.acquire(ctx, lambda: [(k, a.shape, dt) for k, a in ins.items()],
                                                key=("mpc_in", tuple((k, a.shape[0]) for k, a in ins.items())) + lkey)
            d = pin.upload(ins)
            pout = _native.PackedBuffers.acquire(ctx, lambda: out_specs(), key=("mpc_out",) + lkey)
            o = pout.views
        else:
            d = {k: ctx.to_device(a) for k, a in ins.items()}
            o = {k: ctx.empty(shape, t) for k, shape, t in out_specs()}
        dtC, dtc = d.get("tC", d["C"]), d.get("tc", d["c"])
        dtF = d.get("tF", d["F"]) if dyn == _native.DYN_LINEAR else None
        dtf = d.get("tf") if dyn == _native.DYN_LINEAR else None
        coupling = resolve_coupling(self.coupling, B, n, m)

        def launch(cpl):
            ctx.mpc_step_forward(dt, T, B, n, m, d["C"], d["c"], d["F"], F_hat.shape[0], d.get("f"), d["x_nom"], d["u_nom"],
                                 d["lo"], d["hi"], dtC, dtc, dyn, dtF, dtf, params, self.ls_decay,
                                 MAX_LS_TRIALS, self.need_expand,
                                 _native.COUPLING_BATCH if cpl == "batch" else _native.COUPLING_ELEMENT,
                                 o["x"], o["u"], o["Ks"], o["ks"], o["u_first"], o["objs"], o["costs"], o["old"],
                                 o["alphas"], o["n_qp"], o["free"], o["n_ls"], o["flags"])
        try:
            launch(coupling)
        except _native.DiffMpcError as ex:
            # 'auto' guessed that the batch is resident at once (one CTA / one cluster); the launcher knows better
            if not ((self.coupling or DEFAULT_COUPLING) == "auto" and coupling == "batch" and "unsupported" in str(ex)):
                raise
            coupling = "element"
            launch(coupling)
        if packed:
            r = pout.download()
            pin.release(); pout.release()
        else:
            r = {k: v.download() for k, v in o.items()}
        flags = r["flags"]
        if (flags & _native.FLAG_QP_NOT_CONVERGED).any():
            warnings.warn("Projected Newton Quadratic Programming warning: Did not converge")
        if (flags & _native.FLAG_LS_CAPPED).any():
            warnings.warn("MPCstep line search hit the %d-trial cap on %d elements" % (MAX_LS_TRIALS, int((flags & 4).astype(bool).sum())))
        x, u = r["x"], r["u"]
        assert not (np.isnan(np.min(x)) or np.isnan(np.min(u)))     # reference :284-285
        self.back_out = LqrBackOut(n_total_qp_iter=int(r["n_qp"].max(axis=1).sum()))
        u_first = r["u_first"]
        self.for_out = LqrForOut(r["objs"], lambda: scrambled_norm(u_nom - u_first, B, T, m),
                                 lambda: scrambled_norm(u_nom - u, B, T, m), np.mean(r["alphas"]), r["costs"])
        self.aux = dict(Ks=r["Ks"], ks=r["ks"], alphas=r["alphas"], free=r["free"], n_qp=r["n_qp"], n_ls=r["n_ls"],
                        old_costs=r["old"], coupling=coupling, u_first=r["u_first"])
        return x, u

    # ---- plugin path: Python-callable true cost / dynamics ------------------------------------
    def _is_plugin(self):
        if not isinstance(self.true_cost, QuadCost):
            if not callable(self.true_cost):
                raise TypeError("true_cost must be a util.QuadCost or a callable tau[B,s] -> cost[B]")
            return True
        if isinstance(self.true_dynamics, LinDx) or is_pendulum(self.true_dynamics):
            return False
        if not callable(self.true_dynamics):
            raise TypeError("true_dynamics must be a util.LinDx, a PendulumDx or a callable (x[B,n], u[B,m]) -> x_next[B,n]")
        return True

    def _stage_cost(self, t, tau):
        if isinstance(self.true_cost, QuadCost):                     # reference :245-251
            C, c = np.asarray(to_xp(self.true_cost.C)), np.asarray(to_xp(self.true_cost.c))
            return 0.5 * np.einsum("bi,bij,bj->b", tau, C[t], tau) + np.einsum("bi,bi->b", tau, c[t])
        return np.asarray(to_xp(self.true_cost(tau)), dtype=np.float64)

    def _true_step(self, t, x, u):
        if isinstance(self.true_dynamics, LinDx):                    # reference :229-236
            Fm, f = np.asarray(to_xp(self.true_dynamics.F)), to_xp(self.true_dynamics.f)
            nx = np.einsum("bij,bj->bi", Fm[t], np.concatenate((x, u), axis=1))
            return nx if f is None else nx + np.asarray(f)[t]
        return np.asarray(to_xp(self.true_dynamics(x, u)), dtype=np.float64)      # :237-240

    def _host_line_search(self, Ks, ks, x_nom, u_nom, lo, hi):
        """forward_rec (reference :175-286) for callables: per-element alpha, every pass re-rolls the horizon through
        the callable for the whole batch; elements whose cost already dropped keep their alpha, so their rows repeat."""
        T, B, m = self.T, self.n_batch, self.n_ctrl
        old = sum(self._stage_cost(t, np.concatenate((x_nom[t], u_nom[t]), axis=1)) for t in range(T))
        alphas = np.ones(B)
        u_first = None
        n_ls = np.ones(B, dtype=np.int32)               # passes each element needed (the kernel's d_n_ls)
        capped = np.zeros(B, dtype=bool)
        trial = 0
        while True:
            xs, us, objs = [x_nom[0]], [], []
            dx = np.zeros_like(x_nom[0])
            for t in range(T):
                ut = np.einsum("bij,bj->bi", Ks[t], dx) + u_nom[t] + alphas[:, None] * ks[t]
                assert np.isfinite(ut).all()
                ut = np.minimum(np.maximum(ut, lo[t]), hi[t])
                us.append(ut)
                objs.append(self._stage_cost(t, np.concatenate((xs[t], ut), axis=1)))
                if t < T - 1:
                    nx = self._true_step(t, xs[t], ut)
                    assert not np.isnan(nx).any()
                    xs.append(nx)
                    dx = nx - x_nom[t + 1]
            cur = np.sum(np.stack(objs), axis=0)
            if u_first is None:
                u_first = np.stack(us)
            worse = cur > old
            trial += 1
            if not worse.any():
                break
            if trial >= MAX_LS_TRIALS:          # same cap as the kernel: alpha stays the one of the last pass
                capped = worse
                break
            alphas[worse] *= self.ls_decay
            n_ls[worse] += 1
        return np.stack(xs), np.stack(us), np.stack(objs), cur, old, alphas, u_first, n_ls, capped

    def _forward_plugin(self, C_hat, c_hat, F_hat, f_hat, x_nom, u_nom, lo, hi):
        T, B, n, m = self.T, self.n_batch, self.n_state, self.n_ctrl
        ctx, dt = self._ctx, C_hat.dtype
        ins = {"C": C_hat, "c": c_hat, "F": F_hat, "x_nom": x_nom, "u_nom": u_nom, "lo": lo, "hi": hi}
        if not (f_hat is None or self.need_expand):
            ins["f"] = f_hat[:T - 1]
        d = {k: ctx.to_device(a) for k, a in ins.items()}
        o = {k: ctx.empty(shape, t) for k, shape, t in
             [("Ks", (T, B, m, n), dt), ("ks", (T, B, m), dt), ("n_qp", (T, B), np.int32), ("free", (T, B, m), np.uint8),
              ("flags", (B,), np.int32)]}
        coupling = resolve_coupling(self.coupling, B, n, m)

        def launch(cpl):
            ctx.mpc_step_forward(dt, T, B, n, m, d["C"], d["c"], d["F"], F_hat.shape[0], d.get("f"), d["x_nom"], d["u_nom"],
                                 d["lo"], d["hi"], None, None, _native.DYN_LINEAR, None, None, None, self.ls_decay,
                                 -1, self.need_expand,
                                 _native.COUPLING_BATCH if cpl == "batch" else _native.COUPLING_ELEMENT,
                                 None, None, o["Ks"], o["ks"], None, None, None, None, None, o["n_qp"], o["free"], None,
                                 o["flags"])
        try:
            launch(coupling)
        except _native.DiffMpcError as ex:
            if not ((self.coupling or DEFAULT_COUPLING) == "auto" and coupling == "batch" and "unsupported" in str(ex)):
                raise
            coupling = "element"
            launch(coupling)
        r = {k: v.download() for k, v in o.items()}
        if (r["flags"] & _native.FLAG_QP_NOT_CONVERGED).any():
            warnings.warn("Projected Newton Quadratic Programming warning: Did not converge")
        f64 = np.float64
        x, u, objs, costs, old, alphas, u_first, n_ls, capped = self._host_line_search(
            r["Ks"].astype(f64), r["ks"].astype(f64), x_nom.astype(f64), u_nom.astype(f64), lo.astype(f64), hi.astype(f64))
        if capped.any():
            warnings.warn("MPCstep line search hit the %d-trial cap on %d elements" % (MAX_LS_TRIALS, int(capped.sum())))
        x, u = x.astype(dt), u.astype(dt)
        assert not np.isnan(x).any() and not np.isnan(u).any()
        self.back_out = LqrBackOut(n_total_qp_iter=int(r["n_qp"].max(axis=1).sum()))
        self.for_out = LqrForOut(objs, scrambled_norm(u_nom - u_first, B, T, m), scrambled_norm(u_nom - u, B, T, m),
                                 np.mean(alphas), costs)
        self.aux = dict(Ks=r["Ks"], ks=r["ks"], alphas=alphas, free=r["free"], n_qp=r["n_qp"], n_ls=n_ls, old_costs=old,
                        coupling=coupling, u_first=u_first, plugin=True)
        return x, u

    def forward(self, inputs):
        x_init, C_hat, c_hat, F_hat, f_hat = inputs
        self.retain_inputs((0, 1, 2, 3, 4))
        self._fwd = tuple(to_xp(v) for v in inputs)
        if self.no_op_forward:
            self.retain_outputs((0, 1))
            self._out_xu = (np.asarray(self.current_states), np.asarray(self.controls))
            return self.current_states, self.controls
        x, u = self._forward_arrays(to_xp(C_hat), to_xp(c_hat), to_xp(F_hat), to_xp(f_hat))
        self._out_xu = (x, u)
        self.retain_outputs((0, 1))
        return x, u

    # ---- backward --------------------------------------------------------------------------
    def _device_inputs(self, dt, C_hat, c_hat, F_hat, new_x, new_u, lo, hi):
        """Device copies of the tensors the adjoint reads: the caller's (BoxDDP's device loop leaves them in HBM) when it
        provided them for this dtype, uploads of the retained host arrays otherwise."""
        host = dict(C=C_hat, c=c_hat, F=F_hat, x=new_x, u=new_u, lo=lo, hi=hi)
        cache = self._dev_cache or {}
        out = {}
        for k, a in host.items():
            d = cache.get(k)
            if d is not None and d.dtype == np.dtype(dt) and int(np.prod(d.shape)) >= a.size:
                out[k] = d
            else:
                out[k] = self._ctx.to_device(a)
        return out

    def backward_numpy(self, dl_dx, dl_du):
        T, B, n, m, s = self.T, self.n_batch, self.n_state, self.n_ctrl, self.n_sc
        ctx = self._ctx
        x_init, C_hat, c_hat, F_hat, f_hat = self._fwd
        C_hat = as_f(C_hat)
        dt = C_hat.dtype
        c_hat, F_hat = as_f(c_hat, dt), as_f(F_hat, dt)
        new_x, new_u = as_f(self._out_xu[0], dt), as_f(self._out_xu[1], dt)
        lo, hi = as_f(self.u_lower, dt), as_f(self.u_upper, dt)
        gx = None if dl_dx is None else ctx.to_device(as_f(dl_dx, dt))
        gu = None if dl_du is None else ctx.to_device(as_f(dl_du, dt))
        if dl_dx is not None:
            assert list(np.shape(dl_dx)) == [T, B, n]
        if dl_du is not None:
            assert list(np.shape(dl_du)) == [T, B, m]
        FT = F_hat.shape[0]
        wsK = ctx.empty((T, B, m, n), dt); wsk = ctx.empty((T, B, m), dt); wsd = ctx.empty((T, B, s), dt)
        act = ctx.empty((T, B, m), np.uint8)
        dx0 = ctx.empty((B, n), dt); dC = ctx.empty((T, B, s, s), dt); dc = ctx.empty((T, B, s), dt)
        dF = ctx.empty((FT, B, n, s), dt)
        df = ctx.empty((T - 1, B, n), dt) if (f_hat is not None and T > 1) else None
        dv = self._device_inputs(dt, C_hat, c_hat, F_hat, new_x, new_u, lo, hi)
        ctx.mpc_step_backward(dt, T, B, n, m, dv["C"], dv["c"], dv["F"], FT, dv["x"], dv["u"], dv["lo"], dv["hi"], gx, gu,
                              wsK, wsk, wsd, act, dx0, dC, dc, dF, df)
        self.active_index = act.download().astype(bool)
        return dx0.download(), dC.download(), dc.download(), dF.download(), (None if df is None else df.download())

    def backward_reduced_numpy(self, dl_dx, dl_du):
        """MPCstep.backward with the (T,B)-sum of dC, dc, dF, df fused in (dmpc_mpc_step_backward_reduced): returns
        (dx0 [B,n], sum dC [s,s], sum dc [s], sum dF [n,s], sum df [n]) - the gradients IL_Env.mpc's repeated q, p
        (env_dx/il_env.py:120-129) and MpcNet's A, B (mpc_net.py:78-86) receive."""
        T, B, n, m, s = self.T, self.n_batch, self.n_state, self.n_ctrl, self.n_sc
        ctx = self._ctx
        x_init, C_hat, c_hat, F_hat, f_hat = self._fwd
        C_hat = as_f(C_hat)
        dt = C_hat.dtype
        c_hat, F_hat = as_f(c_hat, dt), as_f(F_hat, dt)
        new_x, new_u = as_f(self._out_xu[0], dt), as_f(self._out_xu[1], dt)
        lo, hi = as_f(self.u_lower, dt), as_f(self.u_upper, dt)
        gx = None if dl_dx is None else ctx.to_device(as_f(dl_dx, dt))
        gu = None if dl_du is None else ctx.to_device(as_f(dl_du, dt))
        rsz = ctx.reduced_grad_elems(n, m)
        wsK = ctx.empty((T, B, m, n), dt); wsk = ctx.empty((T, B, m), dt); wsd = ctx.empty((T, B, s), dt)
        act = ctx.empty((T, B, m), np.uint8)
        part = ctx.empty((B, rsz), dt); sums = ctx.empty((rsz,), dt); dx0 = ctx.empty((B, n), dt)
        dv = self._device_inputs(dt, C_hat, c_hat, F_hat, new_x, new_u, lo, hi)
        ctx.mpc_step_backward_reduced(dt, T, B, n, m, dv["C"], dv["c"], dv["F"], F_hat.shape[0], dv["x"], dv["u"], dv["lo"],
                                      dv["hi"], gx, gu, wsK, wsk, wsd, act, part, dx0, sums)
        self.active_index = act.download().astype(bool)
        return (dx0.download(),) + _native.Context.split_reduced(sums.download(), n, m)

    def backward(self, target_input_indexes, grad_outputs):
        dl_dx, dl_du = grad_outputs
        g = self.backward_numpy(to_xp(dl_dx), to_xp(dl_du))
        return tuple(wrap(v) for v in g)

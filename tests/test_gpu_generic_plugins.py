"""GPU: Python-callable true cost / dynamics (reference mpc_step.py:237-251, box_ddp.py:123-136, approximate.py).
backward_rec runs on the GPU (dmpc_mpc_step_forward with max_ls_trials < 0 = sweep only), the line search runs on the
host through the callable.  Checked against the fused device path on the same problem expressed as QuadCost / LinDx /
pendulum, and against the oracle."""
import contextlib
import io
import warnings

import numpy as np
import pytest

from _helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu


def arr(v):
    return np.asarray(getattr(v, "array", v))


class _T:
    """Stage index for callables that stand in for time-varying LinDx / QuadCost (the line search calls them in t order)."""
    def __init__(self, T, per_pass):
        self.T, self.k, self.per = T, 0, per_pass

    def next(self):
        t = self.k % self.per
        self.k += 1
        return t


@pytest.mark.parametrize("name", ["mpc_n3m2", "mpc_n8m4"])
def test_sweep_only_call_returns_the_same_gains(name):
    """C ABI: max_ls_trials < 0 writes Ks, ks, n_qp, free, flags of the full call and touches nothing else."""
    import _native
    g = load_golden(name + "_batch")
    n, m = int(g["n"]), int(g["m"])
    T, B = g["C"].shape[:2]
    ctx = _native.default_context(0)
    dt = np.float64
    d = {k: ctx.to_device(np.ascontiguousarray(g[k])) for k in ("C", "c", "F", "x_nom", "u_nom", "lower", "upper")}
    df = ctx.to_device(g["f"]) if g.get("f") is not None else None

    def outs():
        return dict(x=ctx.empty((T, B, n), dt), u=ctx.empty((T, B, m), dt), Ks=ctx.empty((T, B, m, n), dt),
                    ks=ctx.empty((T, B, m), dt), uf=ctx.empty((T, B, m), dt), objs=ctx.empty((T, B), dt),
                    costs=ctx.empty((B,), dt), old=ctx.empty((B,), dt), al=ctx.empty((B,), dt),
                    nqp=ctx.empty((T, B), np.int32), free=ctx.empty((T, B, m), np.uint8), nls=ctx.empty((B,), np.int32),
                    fl=ctx.empty((B,), np.int32))
    a, b = outs(), outs()
    ctx.mpc_step_forward(dt, T, B, n, m, d["C"], d["c"], d["F"], g["F"].shape[0], df, d["x_nom"], d["u_nom"], d["lower"],
                         d["upper"], d["C"], d["c"], _native.DYN_LINEAR, d["F"], df, None, 0.2, 64, True,
                         _native.COUPLING_BATCH, a["x"], a["u"], a["Ks"], a["ks"], a["uf"], a["objs"], a["costs"], a["old"],
                         a["al"], a["nqp"], a["free"], a["nls"], a["fl"])
    ctx.mpc_step_forward(dt, T, B, n, m, d["C"], d["c"], d["F"], g["F"].shape[0], df, d["x_nom"], d["u_nom"], d["lower"],
                         d["upper"], None, None, _native.DYN_LINEAR, None, None, None, 0.2, -1, True,
                         _native.COUPLING_BATCH, None, None, b["Ks"], b["ks"], None, None, None, None, None, b["nqp"],
                         b["free"], None, b["fl"])
    for k in ("Ks", "ks", "nqp", "free"):
        assert np.array_equal(a[k].download(), b[k].download()), k
    assert np.array_equal(a["fl"].download() & 3, b["fl"].download() & 3)


@pytest.mark.parametrize("name", ["mpc_n3m2", "mpc_n3m1", "mpc_n8m4", "mpc_n4m2_loose"])
def test_mpcstep_with_callables_equals_the_fused_step(name):
    from mpc_step import MPCstep
    from util import QuadCost, LinDx
    g = load_golden(name + "_batch")
    n, m = int(g["n"]), int(g["m"])
    T, B = g["C"].shape[:2]
    f = g.get("f")
    Fm, C, c = g["F"], g["C"], g["c"]
    kd, kc = _T(T, T - 1), _T(T, T)

    def dyn(x, u):
        t = kd.next()
        nx = np.einsum("bij,bj->bi", Fm[t], np.concatenate((arr(x), arr(u)), axis=1))
        return nx if f is None else nx + f[t]

    def cost(tau):
        t = kc.next()
        tau = arr(tau)
        return 0.5 * np.einsum("bi,bij,bj->b", tau, C[t], tau) + np.einsum("bi,bi->b", tau, c[t])

    for tc_, td_ in ((cost, dyn), (QuadCost(C, c), dyn), (cost, LinDx(Fm, f))):
        kd.k = kc.k = 0
        st = MPCstep(controls=g["u_nom"], T=T, u_upper=g["upper"], u_lower=g["lower"], n_batch=B, n_state=n, n_ctrl=m,
                     current_states=g["x_nom"], true_cost=tc_, true_dynamics=td_, ls_decay=0.2, max_ls_iter=10,
                     need_expand=True)
        x, u = st.apply((g["x0"], C, c, Fm, f))
        assert st.aux.get("plugin") and st.aux["coupling"] == "batch"
        assert rel_err(arr(x), g["x"]) < 1e-10 and rel_err(arr(u), g["u"]) < 1e-10
        assert rel_err(st.for_out.costs, g["costs"]) < 1e-10 and rel_err(st.for_out.objs, g["objs"]) < 1e-10
        assert rel_err(st.for_out.full_du_norm, g["full_du_norm"]) < 1e-10
        assert rel_err(st.for_out.alpha_du_norm, g["alpha_du_norm"]) < 1e-10
        assert abs(st.for_out.mean_alphas - float(g["mean_alphas"])) < 1e-15
        assert st.back_out.n_total_qp_iter == int(g["n_total_qp_iter"])


def _pendulum_problem(B, T=20, seed=0):
    rs = np.random.RandomState(seed)
    th = rs.uniform(-np.pi / 2, np.pi / 2, B)
    x0 = np.stack((np.cos(th), np.sin(th), rs.uniform(-1, 1, B)), axis=1)
    q = np.array([1.0, 1.0, 0.1, 0.001]); p = np.array([-1.0, 0.0, 0.0, 0.0])
    C = np.broadcast_to(np.diag(q), (T, B, 4, 4)).copy()
    c = np.broadcast_to(p, (T, B, 4)).copy()
    return x0, C, c, q, p


@pytest.mark.parametrize("clip", [False, True])
def test_boxddp_with_an_opaque_pendulum_callable_and_callable_cost(monkeypatch, clip):
    """The pendulum handed over as a plain Python function and the quadratic cost as a callable: BoxDDP linearises /
    approximates on the host (finite differences), sweeps on the GPU, line-searches on the host, and follows the fused
    device loop iteration for iteration (fixed number of iterations: the comparison must not depend on where a hard
    swing-up instance stops); the adjoint of the final no-op step agrees too."""
    from box_ddp import BoxDDP
    from pendulum_dx import PendulumDx
    from util import QuadCost
    from oracle import pendulum as pend
    # the pendulum clips its torque at +-2 = the control bounds: inside the feasible set the clip never acts, but it puts a
    # kink exactly where clamped controls sit.  clip=False: the callable is the smooth map (central differences, ~1e-10);
    # clip=True: the kink is in the stencil and approximate._fd_column resolves it to the inclusive one-sided slope, which
    # is what the analytic linearisation and Chainer's F.clip use - first-order accurate there.
    if not clip:
        monkeypatch.setattr(pend, "MAX_TORQUE", 1e9)
    tol = 5e-3 if clip else 1e-4
    B, T = 12, 20
    x0, C, c, q, p = _pendulum_problem(B, T)
    kw = dict(T=T, u_lower=-2.0, u_upper=2.0, n_batch=B, n_state=3, n_ctrl=1, u_init=None, eps=1e-9, max_iter=4,
              line_search_decay=0.2, max_line_search_iter=5, update_dynamics=True, exit_unconverged=False)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fused = BoxDDP(**kw)
        xf, uf, cf = fused((x0, QuadCost(C, c), PendulumDx()))
        plug = BoxDDP(**kw)
        xp_, up_, cp_ = plug((x0, lambda tau: 0.5 * (arr(tau) ** 2) @ q + arr(tau) @ p,
                              lambda x, u: pend.step(arr(x), arr(u))))
    assert plug.info["n_iter"] == fused.info["n_iter"] == 4
    # finite-difference derivatives (~1e-7 relative in the Hessian) through four iLQR iterations
    assert np.abs(arr(up_) - arr(uf)).max() < tol and np.abs(arr(xp_) - arr(xf)).max() < tol
    assert np.abs(cp_ - cf).max() < tol * np.abs(cf).max()
    # gradient path: the final no-op MPCstep carries finite-difference C, c, F, f of the callables
    gu = np.random.RandomState(1).randn(T, B, 1)
    gf = fused.last_step.backward_numpy(None, gu)
    gp = plug.last_step.backward_numpy(None, gu)
    for a, b, k in zip(gp, gf, ("dx0", "dC", "dc", "dF", "df")):
        assert np.isfinite(a).all(), k
        assert np.abs(a - b).max() < 10 * tol * max(np.abs(b).max(), 1e-3), k


def test_boxddp_linear_callable_matches_lindx():
    """Linear dynamics behind a callable: finite differences of a linear map are exact to rounding, so the plugin loop
    (host loop, host line search) follows the device loop on the LinDx problem iteration for iteration."""
    from box_ddp import BoxDDP
    from util import QuadCost, LinDx
    g = load_golden("mpc_n3m2_batch")
    n, m = 3, 2
    T, B = g["C"].shape[:2]
    Fm = np.broadcast_to(g["F"][:1], g["F"].shape).copy()          # time-invariant so that the callable needs no t
    f = None
    kw = dict(T=T, u_lower=g["lower"], u_upper=g["upper"], n_batch=B, n_state=n, n_ctrl=m, u_init=None, eps=1e-6,
              max_iter=30, update_dynamics=True)
    sink = io.StringIO()
    with contextlib.redirect_stdout(sink), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        a = BoxDDP(**kw)
        xa, ua, ca = a((g["x0"], QuadCost(g["C"], g["c"]), LinDx(Fm, f)))
        b = BoxDDP(**kw)
        xb, ub, cb = b((g["x0"], QuadCost(g["C"], g["c"]),
                        lambda x, u: np.einsum("bij,bj->bi", Fm[0], np.concatenate((arr(x), arr(u)), axis=1))))
    assert a.info["status"] == b.info["status"]
    assert np.abs(arr(ua) - arr(ub)).max() < 1e-6 and np.abs(arr(xa) - arr(xb)).max() < 1e-6
    assert np.abs(ca - cb).max() < 1e-8 * max(1.0, np.abs(ca).max())


def test_non_simple_pendulum_takes_the_plugin_path():
    """PendulumDx(simple=False) (damping, gravity bias: env_dx/pendulum.py:88-93) has no device code; it is an ordinary
    callable.  With d = b = 0 it is the simple model, so BoxDDP through the plugin path must follow the fused device loop."""
    from box_ddp import BoxDDP
    from mpc_step import is_pendulum
    from pendulum_dx import PendulumDx
    from util import QuadCost
    B, T = 8, 12
    x0, C, c, q, p = _pendulum_problem(B, T, seed=3)
    full = PendulumDx(simple=False, params=[10.0, 1.0, 1.0, 0.0, 0.0])
    assert not is_pendulum(full) and is_pendulum(PendulumDx())
    kw = dict(T=T, u_lower=-2.0, u_upper=2.0, n_batch=B, n_state=3, n_ctrl=1, u_init=None, eps=1e-9, max_iter=3,
              line_search_decay=0.2, max_line_search_iter=5, update_dynamics=True, exit_unconverged=False)
    with contextlib.redirect_stdout(io.StringIO()), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        xf, uf, cf = BoxDDP(**kw)((x0, QuadCost(C, c), PendulumDx()))
        xg, ug, cg = BoxDDP(**kw)((x0, QuadCost(C, c), full))
    assert np.abs(arr(ug) - arr(uf)).max() < 5e-3 and np.abs(arr(xg) - arr(xf)).max() < 5e-3
    assert np.abs(cg - cf).max() < 5e-3 * np.abs(cf).max()

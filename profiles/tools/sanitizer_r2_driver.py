"""compute-sanitizer driver, round 2: the kernels that are new or changed this round, at small sizes.
  - n=32/m=8 persistent DMMA forward with V emission + two-sweep adjoint (full gradients and fused (T,B)-reduction)
  - thread-per-element MPC step: batch coupling in one CTA (candidate stash), element coupling, batch coupling across a
    thread-block cluster; the sweep-only call (max_ls_trials < 0)
  - group MPC step (n=8, m=4), PNQP across a cluster
  - device-resident BoxDDP loop (pendulum: no per-iteration rollout; LinDx) and the reduced backward of its final step
"""
import contextlib, io, os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g   # noqa: E402  (puts the package on sys.path)
import numpy as np            # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from _helpers import lqr_problem   # noqa: E402
import _native                # noqa: E402
from differentiable_lqr import DiffLqr     # noqa: E402
from mpc_step import MPCstep  # noqa: E402
from pnqp import PNQP         # noqa: E402
from box_ddp import BoxDDP    # noqa: E402
from util import QuadCost, LinDx           # noqa: E402
from pendulum_dx import PendulumDx         # noqa: E402

ctx = _native.default_context(0)
quiet = contextlib.redirect_stdout(io.StringIO())
warnings.simplefilter("ignore")

# ---- config-5 shape, short horizon: forward (V emission), adjoint, reduced adjoint
T, B, n, m = 6, 9, 32, 8
pr = lqr_problem(3, T, B, n, m, with_f=True, sym=False)
layer = DiffLqr(T, B, n, m)
x, u = layer.apply((pr["x0"], pr["C"], pr["c"], pr["F"], pr["f"]))
gx, gu = np.ones((T, B, n)), np.ones((T, B, m))
layer.backward_numpy(gx, gu)
layer.backward_reduced_numpy(gx, gu)

# ---- pendulum steps through the thread-per-element kernel
def pend(Bp, Tp=8):
    rs = np.random.RandomState(Bp)
    th = rs.rand(Bp) * np.pi - np.pi / 2
    x0 = np.stack((np.cos(th), np.sin(th), rs.rand(Bp) * 2 - 1), axis=1)
    dx = PendulumDx(); qv, pv = dx.get_true_obj()
    C = np.repeat(np.repeat(np.diag(qv)[None, None], Tp, 0), Bp, 1); c = np.repeat(np.repeat(pv[None, None], Tp, 0), Bp, 1)
    return x0, C, c, dx

for Bp, cpl in ((64, "batch"), (80, "element"), (700, "batch")):
    Tp = 8
    x0, C, c, dx = pend(Bp, Tp)
    unom = np.zeros((Tp, Bp, 1))
    xd = ctx.empty((Tp, Bp, 3)); Fo = ctx.empty((Tp - 1, Bp, 3, 4)); fo = ctx.empty((Tp - 1, Bp, 3))
    ctx.get_traj(np.float64, Tp, Bp, 3, 1, _native.DYN_PENDULUM, ctx.to_device(x0), ctx.to_device(unom), None, None, (10.0, 1.0, 1.0), xd, Fo, fo)
    st = MPCstep(controls=unom, T=Tp, u_upper=np.full((Tp, Bp, 1), 2.0), u_lower=np.full((Tp, Bp, 1), -2.0), n_batch=Bp, n_state=3,
                 n_ctrl=1, current_states=xd.download(), true_cost=QuadCost(C, c), true_dynamics=dx, ls_decay=0.2, max_ls_iter=5,
                 need_expand=True, coupling=cpl)
    st.apply((x0, C, c, Fo.download(), fo.download()))
    if Bp == 80:   # sweep only (plugin path): the cost as a Python callable
        st2 = MPCstep(controls=unom, T=Tp, u_upper=np.full((Tp, Bp, 1), 2.0), u_lower=np.full((Tp, Bp, 1), -2.0), n_batch=Bp,
                      n_state=3, n_ctrl=1, current_states=xd.download(),
                      true_cost=lambda tau: 0.5 * (np.asarray(tau) ** 2) @ np.diag(C[0, 0]) + np.asarray(tau) @ c[0, 0],
                      true_dynamics=dx, ls_decay=0.2, max_ls_iter=5, need_expand=True, coupling=cpl)
        st2.apply((x0, C, c, Fo.download(), fo.download()))

# ---- group MPC step n=8 m=4 (LinDx), PNQP across a cluster
T, B, n, m = 6, 40, 8, 4
pr = lqr_problem(5, T, B, n, m, with_f=True)
xn = np.zeros((T, B, n)); xn[0] = pr["x0"]
for t in range(T - 1):
    xn[t + 1] = np.einsum("bij,bj->bi", pr["F"][t], np.concatenate((xn[t], np.zeros((B, m))), axis=1)) + pr["f"][t]
st = MPCstep(controls=np.zeros((T, B, m)), T=T, u_upper=np.full((T, B, m), 0.5), u_lower=np.full((T, B, m), -0.5), n_batch=B,
             n_state=n, n_ctrl=m, current_states=xn, true_cost=QuadCost(pr["C"], pr["c"]), true_dynamics=LinDx(pr["F"], pr["f"]),
             ls_decay=0.2, max_ls_iter=5, need_expand=True)
st.apply((pr["x0"], pr["C"], pr["c"], pr["F"], pr["f"]))
rs = np.random.RandomState(7); Bq, mq = 1500, 4
L = rs.randn(Bq, mq, mq); H = L @ L.transpose(0, 2, 1) + np.eye(mq)
PNQP(H, rs.randn(Bq, mq), -0.3 * np.ones((Bq, mq)), 0.3 * np.ones((Bq, mq)), coupling="batch")

# ---- BoxDDP device loop: pendulum (x_new becomes x_nom, Jacobian in boxddp_post_kernel) and LinDx; reduced backward
x0, C, c, dx = pend(48, 10)
s1 = BoxDDP(T=10, u_lower=-2.0, u_upper=2.0, n_batch=48, n_state=3, n_ctrl=1, u_init=None, eps=1e-3, max_iter=19,
            exit_unconverged=False, line_search_decay=0.2, max_line_search_iter=5, update_dynamics=False)
with quiet:
    s1((x0, QuadCost(C, c), dx))
s1.last_step.backward_reduced_numpy(None, np.ones((10, 48, 1)))
T, B, n, m = 6, 24, 3, 2
pr = lqr_problem(11, T, B, n, m, with_f=True)
s2 = BoxDDP(T=T, u_lower=np.full((T, B, m), -0.4), u_upper=np.full((T, B, m), 0.4), n_batch=B, n_state=n, n_ctrl=m, u_init=None,
            eps=1e-6, max_iter=9)
with quiet:
    s2((pr["x0"], QuadCost(pr["C"], pr["c"]), LinDx(pr["F"], pr["f"])))
s2.last_step.backward_numpy(np.ones((T, B, n)), np.ones((T, B, m)))
print("sanitizer r2 driver ok")

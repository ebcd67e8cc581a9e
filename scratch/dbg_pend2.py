import sys, os, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); import conftest
from _helpers import load_golden, rel_err
from oracle import mpc as ompc, pendulum as opend
import _native
import mpc_step
from mpc_step import MPCstep
from box_ddp import BoxDDP
from util import QuadCost
from pendulum_dx import PendulumDx
g = load_golden("pendulum_ddp")
x0 = g["x0"]; Q = g["Q"]; p = g["p"]; T, B = 20, x0.shape[0]
lo = np.full((T,B,1), -2.0); hi = np.full((T,B,1), 2.0)
dyn = ("pendulum", (10.,1.,1.))
ctx = _native.default_context(0)
dx = PendulumDx()
solver = BoxDDP(T=T, u_lower=-2.0, u_upper=2.0, n_batch=B, n_state=3, n_ctrl=1, u_init=None, eps=1e-3, max_iter=500)
u_o = np.zeros((T,B,1)); u_c = u_o.copy()
for it in range(16):
    x_o = ompc.get_traj(x0, u_o, dyn); F_o, f_o = opend.linearize(x0, u_o)
    xo2, uo2, fo, aux = ompc.step_forward(Q, p, F_o, f_o, x_o, u_o, lo, hi, (Q,p), dyn, 0.2, 5, 3, 1, need_expand=True, coupling="batch")
    x_c, F_c, f_c = solver._rollout(ctx, x0, u_c, dx)
    st = MPCstep(controls=u_c, T=T, u_upper=hi, u_lower=lo, n_batch=B, n_state=3, n_ctrl=1, current_states=x_c, true_cost=QuadCost(Q,p), true_dynamics=dx, ls_decay=0.2, max_ls_iter=5, need_expand=True, coupling="batch")
    xs, us = st._forward_arrays(Q, p, F_c, f_c)
    d = np.where(st.aux["alphas"] != fo.alphas)[0]
    print(it, "traj %.1e F %.1e | x %.2e u %.2e n_ls max %d differing %s cuda a %s oracle a %s |ks| %s" % (rel_err(x_c, x_o), rel_err(F_c, F_o), rel_err(xs, xo2), rel_err(us, uo2), st.aux["n_ls"].max(), d, st.aux["alphas"][d], fo.alphas[d], np.abs(aux["ks"][:, d]).max(axis=(0,2)) if len(d) else ""))
    u_o = uo2; u_c = us

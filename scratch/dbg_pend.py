import sys, os, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); import conftest
from _helpers import load_golden, rel_err
from oracle import mpc as ompc, pendulum as opend
import _native
from mpc_step import MPCstep
from util import QuadCost
from pendulum_dx import PendulumDx
warnings.simplefilter("ignore")
g = load_golden("pendulum_ddp")
x0 = g["x0"]; Q = g["Q"]; p = g["p"]; T, B = 20, x0.shape[0]
lo = np.full((T,B,1), -2.0); hi = np.full((T,B,1), 2.0)
dyn = ("pendulum", (10.,1.,1.))
ctx = _native.default_context(0)
dx = PendulumDx()
u_o = np.zeros((T,B,1)); u_c = u_o.copy()
for it in range(16):
    # oracle
    x_o = ompc.get_traj(x0, u_o, dyn); F_o, f_o = opend.linearize(x0, u_o)
    xo2, uo2, fo, aux = ompc.step_forward(Q, p, F_o, f_o, x_o, u_o, lo, hi, (Q,p), dyn, 0.2, 5, 3, 1, need_expand=True, coupling="batch")
    # cuda with SAME inputs as oracle (single-step parity)
    st = MPCstep(controls=u_o, T=T, u_upper=hi, u_lower=lo, n_batch=B, n_state=3, n_ctrl=1, current_states=x_o, true_cost=QuadCost(Q,p), true_dynamics=dx, ls_decay=0.2, max_ls_iter=5, need_expand=True, coupling="batch")
    xs, us = st._forward_arrays(Q, p, F_o, f_o)
    print(it, "single-step: x %.2e u %.2e alphas_eq %s free_eq %s ks %.2e Ks %.2e nqp_eq %s" % (rel_err(xs, xo2), rel_err(us, uo2), np.array_equal(st.aux["alphas"], fo.alphas), np.array_equal(st.aux["free"].astype(float), aux["free"]), rel_err(st.aux["ks"], aux["ks"]), rel_err(st.aux["Ks"], aux["Ks"]), np.array_equal(st.aux["n_qp"], aux["n_qp"])), "full_du max", fo.full_du_norm.max())
    if it in (4, 10):
        d = np.where(st.aux["alphas"] != fo.alphas)[0]
        print("  differing elems", d, "cuda alphas", st.aux["alphas"][d], "oracle", fo.alphas[d], "n_ls", st.aux["n_ls"][d], "oracle n_ls", fo.n_ls)
        print("  cuda cost-old", (st.for_out.costs - st.aux["old_costs"])[d], "oracle cost-old", (fo.costs - ompc.traj_cost(x_o, u_o, (Q,p)))[d])
        print("  max|ks| of those", np.abs(aux["ks"][:, d]).max(axis=(0,2)), "u diff per elem", np.abs(us-uo2)[:, d].max(axis=(0,2)))
    u_o = uo2

import sys, os, warnings, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import __graft_entry__ as g
g.smoke_large_state()
g.smoke_mpc()
import numpy as np
from box_ddp import BoxDDP
from util import QuadCost
from pendulum_dx import PendulumDx
rs = np.random.RandomState(0); B, T = 16, 20
th = rs.rand(B) * np.pi - np.pi / 2
x0 = np.stack((np.cos(th), np.sin(th), rs.rand(B) * 2 - 1), axis=1)
dx = PendulumDx(); qv, pv = dx.get_true_obj()
Q = np.repeat(np.repeat(np.diag(qv)[None, None], T, 0), B, 1); p = np.repeat(np.repeat(pv[None, None], T, 0), B, 1)
s = BoxDDP(T=T, u_lower=dx.lower, u_upper=dx.upper, n_batch=B, n_state=3, n_ctrl=1, u_init=None, eps=dx.mpc_eps, max_iter=6,
           exit_unconverged=False, line_search_decay=dx.linesearch_decay, max_line_search_iter=dx.max_linesearch_iter)
with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
    warnings.simplefilter("ignore")
    x, u, c = s((x0, QuadCost(Q, p), dx))
    gr = s.last_step.backward_reduced_numpy(None, np.ones((T, B, 1)))
print("sanitizer driver ok", float(np.abs(np.asarray(getattr(u, "array", u))).max()))

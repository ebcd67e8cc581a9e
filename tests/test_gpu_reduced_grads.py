"""GPU: fused (T,B)-reduction of the parameter gradients (dmpc_lqr_adjoint_reduced / dmpc_mpc_step_backward_reduced,
SURVEY.md section 8f-2) against the sum of the materialised gradients - the backward of util.expand_time_batch
(reference util.py:361-377)."""
import warnings

import numpy as np
import pytest

from _helpers import rel_err, lqr_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,B,n,m", [(8, 37, 3, 1), (12, 100, 4, 2), (10, 70, 8, 4), (6, 9, 5, 3), (7, 33, 32, 8), (1, 5, 4, 2)])
@pytest.mark.parametrize("strict", [True, False])
def test_difflqr_reduced_equals_sum_of_full(T, B, n, m, strict):
    import differentiable_lqr as dl
    pr = lqr_problem(11, T, B, n, m)
    rs = np.random.RandomState(3)
    gx, gu = rs.randn(T, B, n), rs.randn(T, B, m)
    node = dl.DiffLqr(T, B, n, m, strict_reference=strict)
    node.apply_numpy(pr["x0"], pr["C"], pr["c"], pr["F"], pr["f"])
    dx0, dC, dc, dF, df = node.backward_numpy(gx, gu)
    rx0, sC, sc, sF, sf = node.backward_reduced_numpy(gx, gu)
    # n=32/m=8: the full-tensor backward is the two-sweep adjoint (dx0 = v'_0), the reduced one the three-sweep recursion
    assert np.array_equal(rx0, dx0) if (n, m) != (32, 8) else rel_err(rx0, dx0) < 1e-12
    tol = 1e-12
    assert rel_err(sC, dC.sum(axis=(0, 1))) < tol and rel_err(sc, dc.sum(axis=(0, 1))) < tol
    if T > 1:
        assert rel_err(sF, dF.sum(axis=(0, 1))) < tol and rel_err(sf, df.sum(axis=(0, 1))) < tol
    else:
        assert not sF.any() and not sf.any()
    # deterministic: a second call reproduces the sums bit for bit
    again = node.backward_reduced_numpy(gx, gu)
    assert all(np.array_equal(a, b) for a, b in zip(again, (rx0, sC, sc, sF, sf)))


@pytest.mark.parametrize("n,m,with_T_rows", [(3, 2, False), (4, 2, True), (8, 4, False)])
def test_mpc_step_backward_reduced_equals_sum_of_full(n, m, with_T_rows):
    from mpc_step import MPCstep
    from util import QuadCost, LinDx
    from oracle import mpc as ompc
    T, B = 7, 21
    s = n + m
    rs = np.random.RandomState(2)
    L = 0.3 * rs.randn(T, B, s, s)
    C = L @ np.transpose(L, (0, 1, 3, 2)) + np.eye(s)
    c = rs.randn(T, B, s)
    FT = T if with_T_rows else T - 1
    F = np.repeat(np.concatenate((0.9 * np.eye(n) + 0.05 * rs.randn(B, n, n), rs.randn(B, n, m)), axis=2)[None], FT, axis=0)
    f = 0.1 * rs.randn(T - 1, B, n)
    x0 = rs.randn(B, n)
    u = np.clip(0.2 * rs.randn(T, B, m), -0.3, 0.3)
    lo, hi = np.full((T, B, m), -0.3), np.full((T, B, m), 0.3)
    x_nom = ompc.get_traj(x0, u, ("linear", F[:T - 1], f))
    st = MPCstep(controls=u, T=T, u_upper=hi, u_lower=lo, n_batch=B, n_state=n, n_ctrl=m, current_states=x_nom,
                 true_cost=QuadCost(C, c), true_dynamics=LinDx(F[:T - 1], f), ls_decay=0.2, max_ls_iter=10, need_expand=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        st.forward((x0, C, c, F, f))
    gx, gu = rs.randn(T, B, n), rs.randn(T, B, m)
    dx0, dC, dc, dF, df = st.backward_numpy(gx, gu)
    rx0, sC, sc, sF, sf = st.backward_reduced_numpy(gx, gu)
    assert np.array_equal(rx0, dx0)
    assert rel_err(sC, dC.sum(axis=(0, 1))) < 1e-12 and rel_err(sc, dc.sum(axis=(0, 1))) < 1e-12
    assert rel_err(sF, dF.sum(axis=(0, 1))) < 1e-12 and rel_err(sf, df.sum(axis=(0, 1))) < 1e-12

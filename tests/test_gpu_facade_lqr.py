"""GPU: the reference-facing Python classes (same names / signatures as lqr/lqr_recursion.py and
lqr/differentiable_lqr.py) reproduce the reference's notebooks and fixtures."""
import numpy as np
import pytest

from _helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu


def test_boyd_example_like_the_notebook():
    """examples/Boyd_lqr.py:24-63 call sequence; printed gains examples/Boyd_lqr.ipynb:508-558."""
    from lqr_recursion import LqrRecursion
    g = load_golden("boyd")
    T, n, m = 51, 3, 1
    test = LqrRecursion(g["x0"], g["C"], g["c"], g["F"], None, T, n, m)   # F has T rows, as in the example
    Ks, ks = test.backward()
    assert len(Ks) == T and Ks[0].shape == (1, 1, 3)
    assert np.allclose(np.asarray(Ks[0])[0, 0], [-1.86152282, -1.34921019, -0.35888729], atol=5e-9)
    assert np.allclose(np.asarray(Ks[47])[0, 0], [-1.5, -1.5, -0.5], atol=5e-9)
    x, u = test.solve_recursion()
    assert rel_err(np.asarray(x), g["x"]) < 1e-10 and rel_err(np.asarray(u), g["u"]) < 1e-10
    x2, u2 = test.forward(Ks, ks)
    assert np.array_equal(np.asarray(x2), np.asarray(x))


def test_onevar_example():
    from lqr_recursion import LqrRecursion
    g = load_golden("onevar")
    x, u = LqrRecursion(g["x0"], g["C"], g["c"], g["F"], None, 20, 2, 1).solve_recursion()
    assert rel_err(np.asarray(x), g["x"]) < 1e-10


@pytest.mark.parametrize("name", ["lqr_n4m2", "lqr_n3m1_f", "lqr_n5m3_nonsym", "lqr_n8m4"])
def test_difflqr_apply_backward(name):
    from differentiable_lqr import DiffLqr
    g = load_golden(name)
    T, B = g["C"].shape[:2]
    node = DiffLqr(T, B, int(g["n"]), int(g["m"]))
    x, u = node.apply((g["x0"], g["C"], g["c"], g["F"], g.get("f")))
    assert rel_err(np.asarray(getattr(x, "array", x)), g["x"]) < 1e-10
    grads = node.backward((0, 1, 2, 3, 4), (g["gx"], g["gu"]))
    for a, k in zip(grads, ("dx0", "dC", "dc", "dF", "df")):
        assert rel_err(np.asarray(getattr(a, "array", a)), g[k]) < 1e-10, k


def test_lqrnet_training_trace_on_gpu():
    """examples/LQRnet.ipynb:184-203: 190 RMSprop updates through DiffLqr fwd + dF on the GPU
    reproduce the notebook's printed losses to 6 digits."""
    from differentiable_lqr import DiffLqr
    from lqr_recursion import LqrRecursion
    want = {int(r[0]): (r[1], r[2]) for r in load_golden("lqrnet_trace")["trace"]}
    T, n, m, B = 5, 3, 1, 128
    s = n + m
    np.random.seed(42)
    p = np.random.randn(s)
    A = np.eye(n) + 0.2 * np.random.randn(n, n)
    Bm = np.random.randn(n, m)
    expF = np.broadcast_to(np.concatenate((A, Bm), 1), (T - 1, B, n, s)).copy()
    C = np.broadcast_to(np.eye(s), (T, B, s, s)).copy()
    c = np.broadcast_to(p, (T, B, s)).copy()
    np.random.seed(2)
    LA = np.eye(n) + 0.2 * np.random.randn(n, n)
    LB = np.random.randn(n, m)
    msA, msB = np.zeros_like(LA), np.zeros_like(LB)
    node = DiffLqr(T, B, n, m)
    for i in range(191):
        x0 = np.random.randn(B, n)
        xt, ut = [np.asarray(getattr(v, "array", v)) for v in LqrRecursion(x0, C, c, expF, None, T, n, m).solve_recursion()]
        F = np.broadcast_to(np.concatenate((LA, LB), 1), (T - 1, B, n, s)).copy()
        xp_, up_ = node.apply_numpy(x0, C, c, F, None)
        loss = np.mean((ut - up_) ** 2) + np.mean((xt - xp_) ** 2)
        dF = node.backward_numpy(-2 * (xt - xp_) / xp_.size, -2 * (ut - up_) / up_.size)[3]
        gsum = dF.sum(axis=(0, 1))
        for P, G, ms in ((LA, gsum[:, :n], msA), (LB, gsum[:, n:], msB)):
            ms *= 0.99
            ms += 0.01 * G * G
            P -= 1e-2 * G / (np.sqrt(ms) + 1e-8)
        ml = np.mean((LA - A) ** 2) + np.mean((LB - Bm) ** 2)
        if i in want:
            assert abs(loss - want[i][0]) < 1e-6 and abs(ml - want[i][1]) < 1e-6, (i, loss, ml)


@pytest.mark.parametrize("T,B,n,m", [(9, 37, 4, 2), (12, 10, 32, 8), (1, 5, 3, 1)])
def test_difflqr_shared_parameter_entry(T, B, n, m):
    """DiffLqr.apply_shared_numpy: ONE C, c, A|B, f block broadcast over [T,B] on the device (dmpc_expand_time_batch =
    util.expand_time_batch, reference util.py:361-377, as LqrNet_cost_dx.forward uses it, differentiable_lqr.py:237-248)
    gives bit-identical x, u to the host-side broadcast, and backward_reduced_numpy the sum of the full gradients."""
    import differentiable_lqr as dl
    rs = np.random.RandomState(T + n)
    s = n + m
    L = 0.3 * rs.randn(s, s)
    C = L @ L.T + np.eye(s) + 0.05 * rs.randn(s, s)            # not symmetric (Q10)
    c = rs.randn(s)
    A = np.eye(n) * 0.9 + 0.05 * rs.randn(n, n)
    F = np.concatenate((A, rs.randn(n, m)), axis=1)
    f = 0.1 * rs.randn(n)
    x0 = rs.randn(B, n)
    gx, gu = rs.randn(T, B, n), rs.randn(T, B, m)
    bc = lambda a, t: np.ascontiguousarray(np.broadcast_to(a, (t, B) + a.shape))
    node = dl.DiffLqr(T, B, n, m)
    x1, u1 = node.apply_numpy(x0, bc(C, T), bc(c, T), bc(F, T - 1), bc(f, T - 1))
    full = node.backward_numpy(gx, gu)
    node2 = dl.DiffLqr(T, B, n, m)
    x2, u2 = node2.apply_shared_numpy(x0, C, c, F, f)
    assert np.array_equal(x1, x2) and np.array_equal(u1, u2)
    dx0, sC, sc, sF, sf = node2.backward_reduced_numpy(gx, gu)
    assert rel_err(dx0, full[0]) < 1e-12
    assert rel_err(sC, full[1].sum(axis=(0, 1))) < 1e-12 and rel_err(sc, full[2].sum(axis=(0, 1))) < 1e-12
    if T > 1:
        assert rel_err(sF, full[3].sum(axis=(0, 1))) < 1e-12 and rel_err(sf, full[4].sum(axis=(0, 1))) < 1e-12

"""CPU: the numpy oracle reproduces every committed golden fixture.

Fixtures were produced by tests/golden/make_golden.py from the UNMODIFIED
reference (lqr/lqr_recursion.py, lqr/differentiable_lqr.py, mpc/pnqp.py,
mpc/mpc_step.py, mpc/active_constrained_lqr.py, mpc/box_ddp.py,
env_dx/pendulum.py) plus the printed notebook outputs.
"""
import glob
import os
import warnings

import numpy as np
import pytest

from oracle import lqr as olqr, pnqp as opnqp, mpc as ompc, boxddp as obox, pendulum as opend

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    d = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    return {k: d[k] for k in d.files}


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) / max(1.0, float(np.max(np.abs(b)))) if np.size(b) else 0.0


LQR_CASES = ["boyd", "onevar", "lqr_n4m2", "lqr_n4m2_nof", "lqr_n3m1_f", "lqr_n5m3_nonsym", "lqr_n8m4", "lqr_T1"]


@pytest.mark.parametrize("name", LQR_CASES)
def test_lqr_solve(name):
    g = load(name)
    n, m = int(g["n"]), int(g["m"])
    x, u, Ks, ks = olqr.lqr_solve(g["x0"], g["C"], g["c"], g["F"], g.get("f"), n, m)
    assert rel(x, g["x"]) < 1e-12 and rel(u, g["u"]) < 1e-12
    assert rel(Ks, g["Ks"]) < 1e-12 and rel(ks, g["ks"]) < 1e-12


def test_boyd_printed_digits():
    g = load("boyd")   # examples/Boyd_lqr.ipynb:508-558
    assert np.allclose(g["Ks"][0, 0, 0], [-1.86152282, -1.34921019, -0.35888729], atol=5e-9)
    assert np.allclose(g["Ks"][47, 0, 0], [-1.5, -1.5, -0.5], atol=5e-9)


@pytest.mark.parametrize("name", [c for c in LQR_CASES if c not in ("boyd", "onevar", "lqr_T1")])
def test_difflqr_backward(name):
    g = load(name)
    n, m = int(g["n"]), int(g["m"])
    out = olqr.difflqr_backward(g["x0"], g["C"], g["c"], g["F"], g["x"], g["u"], g["gx"], g["gu"], n, m)
    for a, k in zip(out, ("dx0", "dC", "dc", "dF", "df")):
        assert rel(a, g[k]) < 1e-12, k


PNQP_CASES = ["pnqp_kat", "pnqp_d4", "pnqp_d4_warm", "pnqp_d1", "pnqp_d8_loose", "pnqp_d3"]


@pytest.mark.parametrize("name", PNQP_CASES)
@pytest.mark.parametrize("fp32", [False, True])
def test_pnqp(name, fp32):
    g = load(name + ("_fp32lu" if fp32 else ""))
    xi = g.get("x_init")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        x, fac, free, it = opnqp.pnqp(g["H"], g["q"], g["lower"], g["upper"], x_init=xi, lu_fp32=fp32)
        xe, _, fe, ie = opnqp.pnqp(g["H"], g["q"], g["lower"], g["upper"], x_init=xi, lu_fp32=fp32,
                                   coupling="element")
    tol = 1e-6 if fp32 else 1e-12
    assert rel(x, g["x"]) < tol
    assert np.array_equal(free, g["free"]) and it == int(g["it"])
    assert rel(xe, g["x_elem"]) < tol
    assert np.array_equal(fe, g["free_elem"]) and np.array_equal(ie, g["it_elem"])
    if "kat" in g:   # experiment_mpc/Projected_Newton_Quadratic_Programming.py:67-68
        assert np.allclose(x, g["kat"], atol=5e-5)


MPC_CASES = ["mpc_n3m2", "mpc_n3m1", "mpc_n8m4", "mpc_n4m2_loose"]


@pytest.mark.parametrize("name", MPC_CASES)
@pytest.mark.parametrize("coupling", ["batch", "element"])
def test_mpc_step_forward(name, coupling):
    g = load(name + "_" + coupling)
    n, m = int(g["n"]), int(g["m"])
    f = g.get("f")
    x, u, fo, aux = ompc.step_forward(g["C"], g["c"], g["F"], f, g["x_nom"], g["u_nom"], g["lower"], g["upper"],
                                      (g["C"], g["c"]), ("linear", g["F"], f), 0.2, 10, n, m,
                                      need_expand=True, coupling=coupling)
    assert rel(x, g["x"]) < 1e-11 and rel(u, g["u"]) < 1e-11
    assert rel(fo.costs, g["costs"]) < 1e-11
    assert np.array_equal(aux["free"], g["free"])


@pytest.mark.parametrize("name", MPC_CASES)
def test_mpc_step_backward(name):
    g = load(name + "_batch")
    n, m = int(g["n"]), int(g["m"])
    out = ompc.step_backward(g["C"], g["c"], g["F"], g.get("f"), g["x"], g["u"], g["lower"], g["upper"],
                             g["gx"], g["gu"], n, m)
    for a, k in zip(out, ("dx0", "dC", "dc", "dF", "df")):
        if a is None:
            assert k not in g
            continue
        assert rel(a, g[k]) < 1e-11, k


@pytest.mark.parametrize("name", ["ddp_n3m2", "ddp_n3m1"])
def test_boxddp_linear(name):
    g = load(name)
    n, m = int(g["n"]), int(g["m"])
    T = g["C"].shape[0]
    b = float(g["bound"])
    o = obox.box_ddp(g["x0"], (g["C"], g["c"]), ("linear", g["F"], g["f"]), T, -b, b, n, m, eps=1e-7, max_iter=30)
    assert rel(o["x"], g["x"]) < 1e-10 and rel(o["u"], g["u"]) < 1e-10
    assert o["n_iter"] == int(g["n_iter"])


def test_pendulum_step_and_jacobian():
    g = load("pendulum_step")   # env_dx/pendulum.py:65-102
    assert rel(opend.step(g["x"], g["u"]), g["xn"]) < 1e-14
    xn, R, S = opend.jacobian(g["x"], g["u"])
    assert rel(R, g["R"]) < 1e-14 and rel(S, g["S"]) < 1e-14


def test_pendulum_boxddp():
    g = load("pendulum_ddp")    # il_env.py:104-158 wiring
    o = obox.box_ddp(g["x0"], (g["Q"], g["p"]), ("pendulum", (10.0, 1.0, 1.0)), 20, -2.0, 2.0, 3, 1,
                     eps=1e-3, max_iter=500, ls_decay=0.2, max_ls_iter=5)
    assert rel(o["x"], g["x"]) < 1e-9 and rel(o["u"], g["u"]) < 1e-9
    assert o["status"] == "converged" and o["n_iter"] == int(g["n_iter"])


def test_lqrnet_training_trace():
    """examples/LQRnet.ipynb cells 2-10 -> printed trace :184-203 (6 digits): pins DiffLqr
    forward and the dF gradient through 190 compounding RMSprop updates."""
    g = load("lqrnet_trace")["trace"]
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
    want = {int(r[0]): (r[1], r[2]) for r in g}
    for i in range(191):
        x0 = np.random.randn(B, n)
        xt, ut, _, _ = olqr.lqr_solve(x0, C, c, expF, None, n, m)
        F = np.broadcast_to(np.concatenate((LA, LB), 1), (T - 1, B, n, s)).copy()
        xp_, up_, _, _ = olqr.lqr_solve(x0, C, c, F, None, n, m)
        loss = np.mean((ut - up_) ** 2) + np.mean((xt - xp_) ** 2)
        gx = -2 * (xt - xp_) / xp_.size
        gu = -2 * (ut - up_) / up_.size
        dF = olqr.difflqr_backward(x0, C, c, F, xp_, up_, gx, gu, n, m)[3]
        gsum = dF.sum(axis=(0, 1))
        for P, G, ms in ((LA, gsum[:, :n], msA), (LB, gsum[:, n:], msB)):
            ms *= 0.99
            ms += 0.01 * G * G
            P -= 1e-2 * G / (np.sqrt(ms) + 1e-8)
        ml = np.mean((LA - A) ** 2) + np.mean((LB - Bm) ** 2)
        if i in want:
            assert abs(loss - want[i][0]) < 1e-6 and abs(ml - want[i][1]) < 1e-6, (i, loss, ml)

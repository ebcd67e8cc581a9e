"""GPU: size-independent properties at BASELINE-sized batches (where the oracle would take minutes) and
finite-difference checks of the adjoint kernels."""
import numpy as np
import pytest
import torch

import _native
from _helpers import rel_err, lqr_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return _native.default_context(0)


def _solve(ctx, pr, flags=_native.LQR_FACTOR | _native.LQR_ROLLOUT | _native.LQR_SAVE_FAC):
    T, B, n, m = pr["T"], pr["B"], pr["n"], pr["m"]
    d = {k: ctx.to_device(pr[k]) for k in ("x0", "C", "c", "F", "f")}
    o = dict(x=ctx.empty((T, B, n)), u=ctx.empty((T, B, m)), Ks=ctx.empty((T, B, m, n)), ks=ctx.empty((T, B, m)),
             fac=ctx.empty((ctx.lqr_fac_elems(T, B, n, m),)))
    ctx.lqr_solve(np.float64, T, B, n, m, d["x0"], d["C"], d["c"], d["F"], T - 1, d["f"], o["x"], o["u"], o["Ks"],
                  o["ks"], o["fac"], flags)
    ctx.sync()
    return d, o


def _loss_and_grads(ctx, pr, wx, wu, strict):
    """loss = <wx, x> + <wu, u>; returns loss and the adjoint gradients."""
    T, B, n, m = pr["T"], pr["B"], pr["n"], pr["m"]
    s = n + m
    d, o = _solve(ctx, pr)
    x, u = o["x"].download(), o["u"].download()
    g = dict(dx0=ctx.empty((B, n)), dC=ctx.empty((T, B, s, s)), dc=ctx.empty((T, B, s)), dF=ctx.empty((T - 1, B, n, s)),
             df=ctx.empty((T - 1, B, n)))
    ctx.lqr_adjoint(np.float64, T, B, n, m, d["C"], d["c"], d["F"], o["x"], o["u"], ctx.to_device(wx), ctx.to_device(wu),
                    o["Ks"], o["fac"], g["dx0"], g["dC"], g["dc"], g["dF"], g["df"],
                    _native.ADJ_STRICT_REFERENCE if strict else 0)
    ctx.sync()
    return float(np.sum(wx * x) + np.sum(wu * u)), {k: v.download() for k, v in g.items()}


@pytest.mark.parametrize("T,B,n,m", [(6, 3, 4, 2), (5, 2, 32, 8), (7, 4, 3, 1)])
def test_adjoint_matches_finite_differences(ctx, T, B, n, m):
    """Corrected-gradient mode (strict_reference=False): dx0, dc, dF, df equal central finite differences of
    the forward kernel; dC[i,j]+dC[j,i] equals the directional derivative along symmetric perturbations."""
    pr = lqr_problem(21 + n, T, B, n, m)
    rs = np.random.RandomState(2)
    wx, wu = rs.randn(T, B, n), rs.randn(T, B, m)
    _, g = _loss_and_grads(ctx, pr, wx, wu, strict=False)
    eps = 1e-6
    for key, gname in (("x0", "dx0"), ("c", "dc"), ("F", "dF"), ("f", "df")):
        dirn = rs.randn(*pr[key].shape)
        pp, pm_ = dict(pr), dict(pr)
        pp[key] = pr[key] + eps * dirn
        pm_[key] = pr[key] - eps * dirn
        lp, _ = _loss_and_grads(ctx, pp, wx, wu, False)
        lm, _ = _loss_and_grads(ctx, pm_, wx, wu, False)
        fd = (lp - lm) / (2 * eps)
        an = float(np.sum(g[gname] * dirn))
        assert abs(fd - an) <= 2e-6 * max(1.0, abs(fd)), (key, fd, an)
    dirn = rs.randn(*pr["C"].shape)
    dirn = dirn + np.transpose(dirn, (0, 1, 3, 2))
    pp, pm_ = dict(pr), dict(pr)
    pp["C"] = pr["C"] + eps * dirn
    pm_["C"] = pr["C"] - eps * dirn
    lp, _ = _loss_and_grads(ctx, pp, wx, wu, False)
    lm, _ = _loss_and_grads(ctx, pm_, wx, wu, False)
    fd = (lp - lm) / (2 * eps)
    an = float(np.sum(g["dC"] * dirn))
    assert abs(fd - an) <= 2e-6 * max(1.0, abs(fd)), ("C", fd, an)


def test_strict_reference_quirks_relation(ctx):
    """Q1/Q2 of SURVEY.md: strict dC = 0.5 dtau x tau + tau x dtau and df shifted by one step."""
    pr = lqr_problem(5, 6, 5, 4, 2)
    rs = np.random.RandomState(3)
    wx, wu = rs.randn(6, 5, 4), rs.randn(6, 5, 2)
    _, gs = _loss_and_grads(ctx, pr, wx, wu, True)
    _, gc = _loss_and_grads(ctx, pr, wx, wu, False)
    assert rel_err(gs["dC"] + np.transpose(gs["dC"], (0, 1, 3, 2)), 1.5 * (gc["dC"] + np.transpose(gc["dC"], (0, 1, 3, 2)))) < 1e-12
    assert np.array_equal(gs["dF"], gc["dF"]) and np.array_equal(gs["dx0"], gc["dx0"])
    assert rel_err(gs["df"][1:], gc["df"][:-1]) < 1e-14        # df_strict[t+1] == df_correct[t]


def test_c5_shape_linearity_and_kkt(ctx):
    """BASELINE config 5 shape (n=32, m=8, T=100) at B=512 through the DMMA kernel: affine in (x0, c, f) and
    stationarity of the Lagrangian: C_t tau_t + c_t + F_t^T lam_{t+1} - [lam_t; 0] = 0 with the kernel's own
    lambda recursion, checked through dF-free identities (dx0 of a zero-gradient problem is zero)."""
    T, B, n, m = 100, 512, 32, 8
    pr = lqr_problem(31, T, B, n, m)
    rs = np.random.RandomState(32)
    pr2 = dict(pr)
    pr2["x0"] = rs.randn(B, n); pr2["c"] = rs.randn(T, B, n + m); pr2["f"] = 0.1 * rs.randn(T - 1, B, n)
    a = 0.37
    pr3 = dict(pr)
    for k in ("x0", "c", "f"):
        pr3[k] = a * pr[k] + (1 - a) * pr2[k]
    o1, o2, o3 = _solve(ctx, pr)[1], _solve(ctx, pr2)[1], _solve(ctx, pr3)[1]
    for k in ("x", "u"):
        mix = a * o1[k].download() + (1 - a) * o2[k].download()
        assert rel_err(o3[k].download(), mix) < 1e-9
    # optimality: u_t minimises the cost-to-go, so perturbing the feed-forward gain increases the total cost
    x, u = o1["x"].download(), o1["u"].download()
    tau = np.concatenate((x, u), axis=2)
    cost = 0.5 * np.einsum("tbi,tbij,tbj->b", tau, pr["C"], tau) + np.einsum("tbi,tbi->b", tau, pr["c"])
    du = 1e-3 * rs.randn(T, B, m)
    u2 = u + du
    x2 = np.empty_like(x); x2[0] = pr["x0"]
    for t in range(T - 1):
        x2[t + 1] = np.einsum("bij,bj->bi", pr["F"][t], np.concatenate((x2[t], u2[t]), axis=1)) + pr["f"][t]
    tau2 = np.concatenate((x2, u2), axis=2)
    cost2 = 0.5 * np.einsum("tbi,tbij,tbj->b", tau2, pr["C"], tau2) + np.einsum("tbi,tbi->b", tau2, pr["c"])
    assert (cost2 >= cost - 1e-9 * np.abs(cost)).all()
    # and the dynamics constraint holds on the returned trajectory
    xn = np.einsum("tbij,tbj->tbi", pr["F"], tau[:-1]) + pr["f"]
    assert rel_err(xn, x[1:]) < 1e-11


def test_c3_full_size_mpc_step_properties(ctx):
    """BASELINE config 3 (n=8, m=4, T=50, B=16384, bounds active on a sizeable fraction of steps):
    feasibility, monotone cost, line-search/active-set consistency and idempotence-at-a-fixed-point."""
    T, B, n, m = 50, 16384, 8, 4
    s = n + m
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(0)
    f64 = torch.float64
    A = 0.9 * torch.eye(n, dtype=f64, device=dev) + 0.05 * torch.randn(B, n, n, dtype=f64, device=dev, generator=g)
    F = torch.cat((A, torch.randn(B, n, m, dtype=f64, device=dev, generator=g)), dim=2)[None].expand(T - 1, B, n, s).contiguous()
    L = 0.3 * torch.randn(B, s, s, dtype=f64, device=dev, generator=g)
    C = (L @ L.transpose(1, 2) + torch.eye(s, dtype=f64, device=dev))[None].expand(T, B, s, s).contiguous()
    c = torch.randn(T, B, s, dtype=f64, device=dev, generator=g)
    f = 0.1 * torch.randn(T - 1, B, n, dtype=f64, device=dev, generator=g)
    x0 = torch.randn(B, n, dtype=f64, device=dev, generator=g)
    bound = 0.6
    u = torch.zeros(T, B, m, dtype=f64, device=dev)
    lo = torch.full((T, B, m), -bound, dtype=f64, device=dev); hi = -lo
    P = lambda t: t.data_ptr()
    st = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    sh = st.cuda_stream
    x = torch.empty(T, B, n, dtype=f64, device=dev)

    def step(u_in):
        torch.cuda.synchronize()        # torch ops (clone) run on the default stream, the kernels on `st`
        ctx.get_traj(np.float64, T, B, n, m, _native.DYN_LINEAR, P(x0), P(u_in), P(F), P(f), None, P(x), None, None, sh)
        o = dict(x=torch.empty_like(x), u=torch.empty_like(u_in), Ks=torch.empty(T, B, m, n, dtype=f64, device=dev),
                 ks=torch.empty(T, B, m, dtype=f64, device=dev), uf=torch.empty_like(u_in),
                 objs=torch.empty(T, B, dtype=f64, device=dev), costs=torch.empty(B, dtype=f64, device=dev),
                 old=torch.empty(B, dtype=f64, device=dev), al=torch.empty(B, dtype=f64, device=dev),
                 nqp=torch.empty(T, B, dtype=torch.int32, device=dev), fr=torch.empty(T, B, m, dtype=torch.uint8, device=dev),
                 nls=torch.empty(B, dtype=torch.int32, device=dev), fl=torch.empty(B, dtype=torch.int32, device=dev))
        ctx.mpc_step_forward(np.float64, T, B, n, m, P(C), P(c), P(F), T - 1, P(f), P(x), P(u_in), P(lo), P(hi), P(C), P(c),
                             _native.DYN_LINEAR, P(F), P(f), None, 0.2, 64, True, _native.COUPLING_ELEMENT, P(o["x"]),
                             P(o["u"]), P(o["Ks"]), P(o["ks"]), P(o["uf"]), P(o["objs"]), P(o["costs"]), P(o["old"]),
                             P(o["al"]), P(o["nqp"]), P(o["fr"]), P(o["nls"]), P(o["fl"]), sh)
        torch.cuda.synchronize()
        return o
    o1 = step(u)
    # the reference warns ("Did not converge") on the rare QP that needs > 20 Newton steps; the kernel flags it
    assert int((o1["fl"] != 0).sum()) <= B // 1000
    assert int((o1["fl"] & _native.FLAG_LS_CAPPED).ne(0).sum()) == 0
    assert bool(((o1["u"] >= -bound) & (o1["u"] <= bound)).all())                      # feasibility
    assert bool((o1["costs"] <= o1["old"]).all())                                       # monotone (Q5)
    assert bool(torch.allclose(o1["objs"].sum(dim=0), o1["costs"], rtol=1e-12, atol=1e-9))
    clamped = (o1["u"].abs() >= bound - 1e-12)
    frac_steps = float(clamped.any(dim=2).double().mean())
    assert 0.05 < frac_steps < 0.95, frac_steps                                         # bounds really are active
    # a clamped control at alpha = 1 is never reported free with a zero step: free mask 0 => k pinned at the bound
    ks_bound = ((o1["ks"] - (lo - u)).abs() < 1e-14) | ((o1["ks"] - (hi - u)).abs() < 1e-14)
    assert bool(ks_bound[o1["fr"] == 0].all())
    # dynamics hold on the returned trajectory
    tau = torch.cat((o1["x"], o1["u"]), dim=2)
    xn = torch.einsum("tbij,tbj->tbi", F, tau[:-1]) + f
    assert float((xn - o1["x"][1:]).abs().max()) < 1e-10
    # iterate: the cost never increases for any element, and the large majority of the batch reaches a fixed
    # point of the step (box-DDP on a convex QP; a few elements keep toggling their active set for longer)
    o = o1
    for _ in range(20):
        o_next = step(o["u"].clone())
        assert bool((o_next["costs"] <= o_next["old"]).all())
        assert bool(torch.allclose(o_next["old"], o["costs"], rtol=1e-10, atol=1e-9))    # old cost == previous new cost
        du = (o_next["u"] - o["u"]).abs().amax(dim=(0, 2))
        o = o_next
    assert float((du < 1e-6).double().mean()) > 0.9


def test_c5_forward_and_adjoint_are_deterministic(ctx):
    """The n=32/m=8 kernels stream their tiles through single-buffered, just-in-time refilled shared memory (bulk copies
    on the async proxy, cp.async, mbarriers).  Any ordering bug there shows up as a rare, non-reproducible corruption of
    one element (round 2 found one: the C row-block refill could overtake the accumulator loads it follows, once per
    ~1e6 element-steps).  The same inputs solved 40 times must give bit-identical gains, trajectories and gradients."""
    T, B, n, m = 100, 1500, 32, 8
    s = n + m
    dev = torch.device("cuda", 0)
    pr = lqr_problem(31, T, B, n, m)
    d = {k: torch.from_numpy(pr[k]).to(dev) for k in ("x0", "C", "c", "F", "f")}
    rs = np.random.RandomState(1)
    gx = torch.from_numpy(rs.randn(T, B, n)).to(dev); gu = torch.from_numpy(rs.randn(T, B, m)).to(dev)
    f64 = torch.float64
    P = lambda t: t.data_ptr()

    def outs():
        return dict(x=torch.empty(T, B, n, dtype=f64, device=dev), u=torch.empty(T, B, m, dtype=f64, device=dev),
                    Ks=torch.empty(T, B, m, n, dtype=f64, device=dev), ks=torch.empty(T, B, m, dtype=f64, device=dev),
                    fac=torch.zeros(ctx.lqr_fac_elems(T, B, n, m), dtype=f64, device=dev),
                    dx0=torch.empty(B, n, dtype=f64, device=dev), dC=torch.empty(T, B, s, s, dtype=f64, device=dev),
                    dc=torch.empty(T, B, s, dtype=f64, device=dev), dF=torch.empty(T - 1, B, n, s, dtype=f64, device=dev),
                    df=torch.empty(T - 1, B, n, dtype=f64, device=dev))
    st = torch.cuda.Stream()

    def run(o):
        ctx.lqr_solve(np.float64, T, B, n, m, P(d["x0"]), P(d["C"]), P(d["c"]), P(d["F"]), T - 1, P(d["f"]), P(o["x"]),
                      P(o["u"]), P(o["Ks"]), P(o["ks"]), P(o["fac"]), _native.LQR_FACTOR | _native.LQR_ROLLOUT | _native.LQR_SAVE_FAC,
                      st.cuda_stream)
        ctx.lqr_adjoint(np.float64, T, B, n, m, P(d["C"]), P(d["c"]), P(d["F"]), P(o["x"]), P(o["u"]), P(gx), P(gu), P(o["Ks"]),
                        P(o["fac"]), P(o["dx0"]), P(o["dC"]), P(o["dc"]), P(o["dF"]), P(o["df"]), _native.ADJ_STRICT_REFERENCE,
                        st.cuda_stream)
        torch.cuda.synchronize()
    ref, o = outs(), outs()
    run(ref)
    for it in range(40):
        run(o)
        for k in ("Ks", "ks", "x", "u", "dx0", "dC", "dc", "dF", "df"):
            assert torch.equal(o[k], ref[k]), (it, k)

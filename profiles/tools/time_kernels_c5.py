"""Times the fused forward launch and the adjoint for config 5 (B=8192) with CUDA events; prints one line."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "chainer-differentiable-mpc_b200"))
import numpy as np, torch
import bench, _native
n, m, T, B = 32, 8, 100, 8192
s = n + m
dev = torch.device("cuda", 0)
ctx = _native.Context(0)
pr = bench.make_problem_torch(torch, dev, n, m, T, B, seed=1)
f64 = torch.float64
o = dict(x=torch.empty(T, B, n, dtype=f64, device=dev), u=torch.empty(T, B, m, dtype=f64, device=dev),
         Ks=torch.empty(T, B, m, n, dtype=f64, device=dev), ks=torch.empty(T, B, m, dtype=f64, device=dev),
         fac=torch.empty(ctx.lqr_fac_elems(T, B, n, m), dtype=f64, device=dev), dx0=torch.empty(B, n, dtype=f64, device=dev),
         dC=torch.empty(T, B, s, s, dtype=f64, device=dev), dc=torch.empty(T, B, s, dtype=f64, device=dev),
         dF=torch.empty(T - 1, B, n, s, dtype=f64, device=dev), df=torch.empty(T - 1, B, n, dtype=f64, device=dev))
P = lambda t: t.data_ptr()
st = torch.cuda.Stream(device=dev)
def fwd(flags):
    ctx.lqr_solve(np.float64, T, B, n, m, P(pr["x0"]), P(pr["C"]), P(pr["c"]), P(pr["F"]), T - 1, P(pr["f"]), P(o["x"]), P(o["u"]),
                  P(o["Ks"]), P(o["ks"]), P(o["fac"]), flags, st.cuda_stream)
def bwd():
    ctx.lqr_adjoint(np.float64, T, B, n, m, P(pr["C"]), P(pr["c"]), P(pr["F"]), P(o["x"]), P(o["u"]), P(pr["gx"]), P(pr["gu"]),
                    P(o["Ks"]), P(o["fac"]), P(o["dx0"]), P(o["dC"]), P(o["dc"]), P(o["dF"]), P(o["df"]), 1, st.cuda_stream)
rsz = ctx.reduced_grad_elems(n, m)
o["part"] = torch.empty(B, rsz, dtype=f64, device=dev); o["sums"] = torch.empty(rsz, dtype=f64, device=dev)
def bwd_red():
    ctx.lqr_adjoint_reduced(np.float64, T, B, n, m, P(pr["C"]), P(pr["c"]), P(pr["F"]), P(o["x"]), P(o["u"]), P(pr["gx"]), P(pr["gu"]),
                            P(o["Ks"]), P(o["fac"]), P(o["dc"]), P(o["part"]), P(o["dx0"]), P(o["sums"]), 1, st.cuda_stream)
FULL = 7
res = {"full": [], "factor": [], "bwd": [], "bwd_reduced": []}
for it in range(7):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    e[0].record(st); fwd(FULL); e[1].record(st); fwd(5); e[2].record(st); bwd(); e[3].record(st); bwd_red(); e[4].record(st)
    torch.cuda.synchronize()
    if it >= 2:
        res["full"].append(e[0].elapsed_time(e[1])); res["factor"].append(e[1].elapsed_time(e[2])); res["bwd"].append(e[2].elapsed_time(e[3])); res["bwd_reduced"].append(e[3].elapsed_time(e[4]))
print(sys.argv[1] if len(sys.argv) > 1 else "", json.dumps({k: round(float(np.median(v)), 3) for k, v in res.items()}), "x_absmax", float(o["x"].abs().max()))

"""Config 5 in fp32 (SURVEY 8d lists c5 as "fp64 and fp32"): times lqr_solve + lqr_adjoint at B=8192 with CUDA
events on the launching stream and checks x, u, dF against the fp64 device result on the same inputs
(north_star tolerance for fp32: 1e-4 relative).  Prints one JSON line."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "chainer-differentiable-mpc_b200"))
import numpy as np, torch
import bench, _native

n, m, T = 32, 8, 100
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
s = n + m
dev = torch.device("cuda", 0)
ctx = _native.Context(0)
st = torch.cuda.Stream(device=dev)
P = lambda t: t.data_ptr()


def outputs(dt):
    return dict(x=torch.empty(T, B, n, dtype=dt, device=dev), u=torch.empty(T, B, m, dtype=dt, device=dev),
                Ks=torch.empty(T, B, m, n, dtype=dt, device=dev), ks=torch.empty(T, B, m, dtype=dt, device=dev),
                fac=torch.empty(ctx.lqr_fac_elems(T, B, n, m), dtype=dt, device=dev), dx0=torch.empty(B, n, dtype=dt, device=dev),
                dC=torch.empty(T, B, s, s, dtype=dt, device=dev), dc=torch.empty(T, B, s, dtype=dt, device=dev),
                dF=torch.empty(T - 1, B, n, s, dtype=dt, device=dev), df=torch.empty(T - 1, B, n, dtype=dt, device=dev))


def fwd(npdt, pr, o, flags=7):
    ctx.lqr_solve(npdt, T, B, n, m, P(pr["x0"]), P(pr["C"]), P(pr["c"]), P(pr["F"]), T - 1, P(pr["f"]), P(o["x"]), P(o["u"]),
                  P(o["Ks"]), P(o["ks"]), P(o["fac"]), flags, st.cuda_stream)


def bwd(npdt, pr, o):
    ctx.lqr_adjoint(npdt, T, B, n, m, P(pr["C"]), P(pr["c"]), P(pr["F"]), P(o["x"]), P(o["u"]), P(pr["gx"]), P(pr["gu"]),
                    P(o["Ks"]), P(o["fac"]), P(o["dx0"]), P(o["dC"]), P(o["dc"]), P(o["dF"]), P(o["df"]), 1, st.cuda_stream)


pr64 = bench.make_problem_torch(torch, dev, n, m, T, B, seed=1)
o64 = outputs(torch.float64)
fwd(np.float64, pr64, o64); bwd(np.float64, pr64, o64)
torch.cuda.synchronize()
ref = {k: o64[k].float() for k in ("x", "u", "dF", "dC")}
del o64
pr32 = {k: v.float() for k, v in pr64.items()}
del pr64
torch.cuda.empty_cache()
o32 = outputs(torch.float32)
res = {"fwd": [], "bwd": [], "factor_only": [], "rollout_only": []}
for it in range(6):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    e[0].record(st); fwd(np.float32, pr32, o32); e[1].record(st); bwd(np.float32, pr32, o32); e[2].record(st)
    fwd(np.float32, pr32, o32, 5); e[3].record(st); fwd(np.float32, pr32, o32, 2); e[4].record(st)   # Riccati sweep alone, rollout alone
    torch.cuda.synchronize()
    if it >= 2:
        res["fwd"].append(e[0].elapsed_time(e[1])); res["bwd"].append(e[1].elapsed_time(e[2]))
        res["factor_only"].append(e[2].elapsed_time(e[3])); res["rollout_only"].append(e[3].elapsed_time(e[4]))
ms = {k: float(np.median(v)) for k, v in res.items()}
rel = {k: float((o32[k] - ref[k]).norm() / ref[k].norm()) for k in ref}
tot = ms["fwd"] + ms["bwd"]
fb, tb = bench.algorithmic_bytes(n, m, T, w=4)
print(json.dumps({"workload": "c5 fp32 (lqr_factor_dmma_warp_kernel<4,float> + adjoint kernels)", "batch": B,
                  "fwd_ms": round(ms["fwd"], 3), "bwd_ms": round(ms["bwd"], 3),
                  "factor_only_ms": round(ms["factor_only"], 3), "rollout_only_ms": round(ms["rollout_only"], 3), "solves_per_sec": B / (tot * 1e-3),
                  "algorithmic_bytes_per_solve": tb, "whole_step_gbs": B * tb / (tot * 1e-3) / 1e9,
                  "rel_err_vs_fp64_device": rel}))

#!/usr/bin/env python
"""Build-container only: time the UNMODIFIED reference (DiffLqr.apply + .backward from /root/reference/lqr, imported
under tests/_chainer_stub) next to the oracle port on the same inputs and host, for BASELINE.md section 3's promise
("the reference's own modules, unmodified, under the stub").  /root/reference does not exist on the GPU box, so bench.py
cannot run this there; it reports the oracle port live and carries this file's ratio as a recorded annotation.

    python profiles/tools/time_reference_under_stub.py > profiles/r2/reference_under_stub.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests", "golden")]
import _ref_loader  # noqa: E402


def main():
    mods = _ref_loader.load(lu_fp32=False)
    import chainer
    V = chainer.Variable
    from oracle import lqr as olqr
    import bench
    out = {"host": {"cpus": os.cpu_count(), "where": "build container (no GPU)"}, "cases": []}
    for (n, m, T, B) in ((32, 8, 100, 128), (4, 2, 50, 4096)):
        pr = bench._Local.lqr_problem_np(0, T, B, n, m)
        rs = np.random.RandomState(1)
        gx, gu = rs.randn(T, B, n), rs.randn(T, B, m)
        DiffLqr = mods["differentiable_lqr"].DiffLqr

        def ref():
            node = DiffLqr(T, B, n, m)
            node.apply((V(pr["x0"]), V(pr["C"]), V(pr["c"]), V(pr["F"]), V(pr["f"])))
            node.backward((0, 1, 2, 3, 4), (V(gx), V(gu)))

        def port():
            x, u, _, _ = olqr.lqr_solve(pr["x0"], pr["C"], pr["c"], pr["F"], pr["f"], n, m)
            olqr.difflqr_backward(pr["x0"], pr["C"], pr["c"], pr["F"], x, u, gx, gu, n, m)
        res = {}
        for name, fn in (("reference_under_stub", ref), ("oracle_port", port)):
            fn()
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
            res[name] = B / min(ts)
        out["cases"].append({"n": n, "m": m, "T": T, "B_cpu": B, "solves_per_sec": res,
                             "port_over_reference": res["oracle_port"] / res["reference_under_stub"]})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

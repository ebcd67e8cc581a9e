#!/usr/bin/env python
"""ncu driver: a few batch-64 pendulum MPC steps (BASELINE config 1) and one config-3 sweep."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (sets up sys.path for the package)
import _native  # noqa: E402
import torch  # noqa: E402

ctx = _native.Context(0)
print(bench.mpc_step_latency(ctx, n_calls=3, cpu_calls=0, with_cpu=False))
if len(sys.argv) > 1 and sys.argv[1] == "c3":
    print(bench.mpc_step_throughput(ctx, torch, torch.device("cuda", 0), reps=1))

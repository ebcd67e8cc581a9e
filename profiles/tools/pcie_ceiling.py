#!/usr/bin/env python
"""Pinned-copy ceiling of the host <-> device feed with N concurrent ranks (one per GPU), for the e2e numbers of bench.py:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        profiles/tools/pcie_ceiling.py

Every rank copies a 512 MB pinned buffer H2D, D2H and both at once (two streams); all ranks start together (barrier),
time = max over ranks.  Printed by rank 0: per-rank and aggregate GB/s per direction, and the e2e solves/s ceiling of
config 5 that follows from it (2.38 MB up + 2.38 MB down per solve through the full-tensor API)."""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 512 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device=dev); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
s1 = torch.cuda.Stream(); s2 = torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d1.copy_(h1, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)


def both():
    h2d(); d2h()


def timed(fn, reps=4):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return reps * n / float(tt.item()) / 1e9


res = {"h2d": timed(h2d), "d2h": timed(d2h), "duplex_per_direction": timed(both)}
if rank == 0:
    per_solve = 2383360.0
    out = {"n_ranks": world, "per_rank_gbs": res, "aggregate_gbs": {k: v * world for k, v in res.items()},
           "config5_e2e_ceiling_solves_per_sec": world * res["duplex_per_direction"] * 1e9 / per_solve,
           "host_cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    print(json.dumps(out))
if world > 1:
    dist.barrier(); dist.destroy_process_group()

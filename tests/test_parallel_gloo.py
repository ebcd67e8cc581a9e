"""CPU, world_size=2 over gloo: batch sharding + parameter-gradient all-reduce.

Each rank solves its contiguous shard (here with the oracle, as no GPU exists in CI), reduces the
per-(t,b) gradients to shared-parameter gradients, and all-reduces them; the result must equal the
full-batch gradients, and the concatenated shard solutions must equal the full-batch solution."""
import os
import socket

import numpy as np
import pytest

import parallel
from _helpers import lqr_problem, rel_err
from oracle import lqr as olqr


def test_shard_bounds_cover_the_batch():
    for B in (1, 7, 64, 65536):
        for world in (1, 2, 3, 8):
            cuts = [parallel.shard_bounds(B, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T, B, n, m = 6, 10, 4, 2
    pr = lqr_problem(3, T, B, n, m)
    rs = np.random.RandomState(4)
    gx, gu = rs.randn(T, B, n), rs.randn(T, B, m)
    lo, hi = parallel.shard_bounds(B, world, rank)
    sh = {k: parallel.shard_tb(pr[k], lo, hi) for k in ("C", "c", "F", "f")}
    x0 = parallel.shard_tb(pr["x0"], lo, hi, batch_axis=0)
    x, u, _, _ = olqr.lqr_solve(x0, sh["C"], sh["c"], sh["F"], sh["f"], n, m)
    g = olqr.difflqr_backward(x0, sh["C"], sh["c"], sh["F"], x, u, parallel.shard_tb(gx, lo, hi),
                              parallel.shard_tb(gu, lo, hi), n, m)
    local = parallel.reduce_param_grads(dC=g[1], dc=g[2], dF=g[3], df=g[4])
    total = parallel.allreduce_param_grads(local)
    q.put((rank, lo, hi, x, u, total))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sharded_solve_and_grad_allreduce_gloo():
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=100) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    T, B, n, m = 6, 10, 4, 2
    pr = lqr_problem(3, T, B, n, m)
    rs = np.random.RandomState(4)
    gx, gu = rs.randn(T, B, n), rs.randn(T, B, m)
    x, u, _, _ = olqr.lqr_solve(pr["x0"], pr["C"], pr["c"], pr["F"], pr["f"], n, m)
    g = olqr.difflqr_backward(pr["x0"], pr["C"], pr["c"], pr["F"], x, u, gx, gu, n, m)
    want = parallel.reduce_param_grads(dC=g[1], dc=g[2], dF=g[3], df=g[4])
    xs = np.concatenate([r[3] for r in res], axis=1)
    us = np.concatenate([r[4] for r in res], axis=1)
    assert np.array_equal(xs, x) and np.array_equal(us, u)       # shard-wise result == full batch
    for r in res:
        for k in want:
            assert rel_err(r[5][k], want[k]) < 1e-12, k

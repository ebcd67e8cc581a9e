"""Imitation learning of the pendulum cost through the differentiable MPC solver, wired as the reference's
env_dx/il_exp.py:230-302 + env_dx/il_env.py:104-158 + env_dx/pendulum_net.py:18-39 wire it, on the B200 path:

    q = sigmoid(q_logit),  p = sqrt(q) * learn_p                       (Pendulum_Net_cost_logit.forward)
    Q, p repeated to [T,B,4,4] / [T,B,4]; BoxDDP(eps 1e-3, max_iter 500, decay 0.2, 5 line-search trials,
    update_dynamics=False, exit_unconverged=False, detach_unconverged=True), warm-started from the last controls  (IL_Env.mpc);
    the warm-start cache of il_exp.py:215-257 (controls per sample id) lives in HBM (_native.WarmStartCache)
    loss = mean((u_expert - u)^2); backward through the final MPCstep; RMSprop(lr 1e-2, alpha 0.5),
    p and q updated in alternating blocks of 10 iterations                                              (IL_Exp.run)

What differs from the reference is only where the work runs: the BoxDDP loop is device resident (dmpc_boxddp_solve), and
the backward of the (T,B)-repeat of q and p - a sum over T and B of dC, dc - is fused into the adjoint kernel
(dmpc_mpc_step_backward_reduced), so the [T,B,4,4] gradient is never materialised.  Under torchrun the batch is sharded
over the ranks and the eight parameter-gradient doubles are all-reduced (NCCL) once per iteration.

    python examples/il_pendulum.py --batch 1024 --iters 60
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/il_pendulum.py --batch 65536
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "chainer-differentiable-mpc_b200")
for p in (PKG, os.path.join(PKG, "lqr"), os.path.join(PKG, "mpc"), os.path.join(PKG, "env_dx")):
    if p not in sys.path:
        sys.path.insert(0, p)

from box_ddp import BoxDDP            # noqa: E402
from pendulum_dx import PendulumDx    # noqa: E402
from util import QuadCost             # noqa: E402
import parallel                       # noqa: E402

T = 20


def mpc(dx, x0, q, p, u_init, device):
    """IL_Env.mpc (il_env.py:104-158): returns (u [T,B,1], solver)."""
    B = x0.shape[0]
    Q = np.broadcast_to(np.diag(q)[None, None], (T, B, 4, 4)).copy()
    pp = np.broadcast_to(p[None, None], (T, B, 4)).copy()
    solver = BoxDDP(T=T, u_lower=dx.lower, u_upper=dx.upper, n_batch=B, n_state=3, n_ctrl=1, u_init=u_init,
                    eps=dx.mpc_eps, max_iter=500, exit_unconverged=False, detach_unconverged=True,
                    line_search_decay=dx.linesearch_decay, max_line_search_iter=dx.max_linesearch_iter,
                    update_dynamics=False, device=device)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter("ignore")
        x, u, _ = solver((x0, QuadCost(Q, pp), dx))
    return np.asarray(getattr(u, "array", u)), solver


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024, help="global batch (sharded over the ranks)")
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    tdev = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        tdev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=tdev)

    dx = PendulumDx()
    true_q, true_p = dx.get_true_obj()
    rs = np.random.RandomState(args.seed)                      # il_env.py:48-70: theta ~ U(-pi/2, pi/2), dtheta ~ U(-1, 1)
    th = rs.uniform(-np.pi / 2, np.pi / 2, args.batch)
    x0_all = np.stack((np.cos(th), np.sin(th), rs.uniform(-1, 1, args.batch)), axis=1)
    lo, hi = parallel.shard_bounds(args.batch, world, rank)
    x0 = np.ascontiguousarray(x0_all[lo:hi])
    B = hi - lo

    u_exp, _ = mpc(dx, x0, np.asarray(true_q, float), np.asarray(true_p, float), None, local)   # expert = true cost

    q_logit, learn_p = np.zeros(4), np.zeros(4)                # pendulum_net.py:22-25
    ms = {"q": np.zeros(4), "p": np.zeros(4)}
    lr, alpha, eps = 1e-2, 0.5, 1e-8                           # chainer.optimizers.RMSprop(lr=1e-2, alpha=0.5), il_exp.py:213
    import _native
    cache = _native.WarmStartCache(_native.default_context(local), B, T, 1)    # il_exp.py:215: zeros, one row per sample id
    ids = np.arange(B)
    update_q = False
    hist = []
    t_start = time.perf_counter()
    for it in range(args.iters):
        if it > 0 and it % 10 == 0:                            # il_exp.py:233-234 round robin
            update_q = not update_q
        q = 1.0 / (1.0 + np.exp(-q_logit))
        p = np.sqrt(q) * learn_p
        u, solver = mpc(dx, x0, q, p, cache.take(ids), local)     # il_exp.py:249: warm start of this minibatch, on the device
        cache.put(ids, solver.u_device)                            # il_exp.py:257, without a host round trip
        diff = u - u_exp
        loss_sum = float(np.sum(diff * diff))
        gu = 2.0 * diff / (args.batch * T)                     # d mean((u_exp-u)^2) / du over the GLOBAL batch
        mask = solver.info.get("detach_mask")
        if mask is not None:                                   # box_ddp.py:263-289: unconverged elements carry no gradient
            gu = gu * mask[None, :, None]
        g = solver.last_step.backward_reduced_numpy(None, gu)  # (dx0, sum dC [4,4], sum dc [4], sum dF, sum df)
        grads = {"q": np.diag(g[1]).copy(), "p": np.asarray(g[2]).copy(), "loss": np.array([loss_sum])}
        grads = parallel.allreduce_param_grads(grads, device=tdev)
        dq, dp = grads["q"], grads["p"]
        loss = float(grads["loss"][0]) / (args.batch * T)
        # chain rule of pendulum_net.py:34-35
        g_logit = (dq + dp * learn_p * 0.5 / np.sqrt(q)) * q * (1.0 - q)
        g_p = dp * np.sqrt(q)
        if update_q:
            ms["q"] = alpha * ms["q"] + (1 - alpha) * g_logit ** 2
            q_logit = q_logit - lr * g_logit / (np.sqrt(ms["q"]) + eps)
        else:
            ms["p"] = alpha * ms["p"] + (1 - alpha) * g_p ** 2
            learn_p = learn_p - lr * g_p / (np.sqrt(ms["p"]) + eps)
        hist.append(loss)
        if rank == 0:
            print("iter %3d  imitation loss %.6e  iLQR iterations %3d  updating %s" %
                  (it, loss, solver.info["n_iter"], "q" if update_q else "p"), file=sys.stderr)
    wall = time.perf_counter() - t_start
    if rank == 0:
        q = 1.0 / (1.0 + np.exp(-q_logit))
        print(json.dumps({"batch": args.batch, "world": world, "iters": args.iters, "loss_first": hist[0], "loss_last": hist[-1],
                          "loss_min": min(hist), "s_per_iter": wall / args.iters,
                          "mpc_solves_per_sec": args.batch * args.iters / wall,
                          "learned_q": q.tolist(), "learned_p": (np.sqrt(q) * learn_p).tolist(),
                          "true_q": np.asarray(true_q).tolist(), "true_p": np.asarray(true_p).tolist()}))
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

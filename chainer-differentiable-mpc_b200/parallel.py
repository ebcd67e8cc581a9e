"""Multi-GPU plumbing: one process per GPU, the batch dimension sharded, no traffic inside a solve.

SURVEY.md section 8(e): batch elements are independent inside every operator, so every [T,B,...]
tensor is split into contiguous B/G slices; the only exchange is one all-reduce(sum) of the
already (T, B_local)-reduced *parameter* gradients per training iteration (e.g. n*s doubles for the
learned A,B of LqrNet, differentiable_lqr.py:170-172; 2*n_sc for il_exp's q, p).  torch.distributed
is the transport (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(B, world, rank):
    """Contiguous [lo, hi) slice of a batch of B for `rank` of `world` (sizes differ by at most 1)."""
    base, rem = divmod(int(B), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_tb(arr, lo, hi, batch_axis=1):
    """Slice a time-major [T,B,...] (or [B,...] with batch_axis=0) array and make it contiguous."""
    if arr is None:
        return None
    idx = [slice(None)] * np.ndim(arr)
    idx[batch_axis] = slice(lo, hi)
    return np.ascontiguousarray(np.asarray(arr)[tuple(idx)])


def reduce_param_grads(dC=None, dc=None, dF=None, df=None):
    """Backward of util.expand_time_batch (util.py:361-377): sum the per-(t,b) gradients over T and
    B_local.  Returns a dict with the shared-parameter gradients present."""
    out = {}
    if dC is not None:
        out["C"] = np.asarray(dC).sum(axis=(0, 1))
    if dc is not None:
        out["c"] = np.asarray(dc).sum(axis=(0, 1))
    if dF is not None:
        out["F"] = np.asarray(dF).sum(axis=(0, 1))
    if df is not None:
        out["f"] = np.asarray(df).sum(axis=(0, 1))
    return out


def allreduce_param_grads(grads, device=None, group=None):
    """One all-reduce(sum) of the packed parameter gradients (a few hundred doubles: latency-bound)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grads
    keys = sorted(grads)
    flat = np.concatenate([np.asarray(grads[k], dtype=np.float64).ravel() for k in keys])
    t = torch.from_numpy(flat)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    flat = t.cpu().numpy()
    out, o = {}, 0
    for k in keys:
        sz = int(np.size(grads[k]))
        out[k] = flat[o:o + sz].reshape(np.shape(grads[k]))
        o += sz
    return out

"""Shot sharding and counter reduction for multi-GPU runs.

Shots are independent, so the path shards with no data-path collective: rank g of G takes the
global shot indices [g*B/G, (g+1)*B/G) and the Philox subsequence is the global shot index, which
makes the sampled error set independent of G.  The only exchange is one all-reduce of the
int64 counter vector at the end of a run (sum, except slot 7 = min_logical_weight which is a min).
Backend-agnostic: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np

COUNTER_NAMES = ("shots", "bp_converged", "bp_success", "osd0_success", "osdw_success",
                 "osd_invocations", "bp_iterations", "min_logical_weight")
MIN_SLOT = 7


def shard_range(total: int, rank: int, world: int):
    """Contiguous, balanced partition of range(total): returns (start, count) of `rank`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("rank/world out of range")
    base, rem = divmod(int(total), world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


def merge_counters(a, b) -> np.ndarray:
    """Combine two counter vectors on the host (same rule the collective applies)."""
    a = np.asarray(a, dtype=np.int64).copy()
    b = np.asarray(b, dtype=np.int64)
    m = _min_pos(a[MIN_SLOT], b[MIN_SLOT])
    a += b
    a[MIN_SLOT] = m
    return a


def _min_pos(x, y):
    # 0 means "no logical failure seen yet"
    if x <= 0:
        return y
    if y <= 0:
        return x
    return min(x, y)


def all_reduce_counters(counters, group=None, device=None) -> np.ndarray:
    """One all-reduce(SUM) over the counter vector; the min slot rides along as a second tiny MIN reduce."""
    import torch
    import torch.distributed as dist

    c = np.asarray(counters, dtype=np.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return c.copy()
    t = torch.from_numpy(c.copy())
    big = np.iinfo(np.int64).max
    mn = torch.tensor([c[MIN_SLOT] if c[MIN_SLOT] > 0 else big], dtype=torch.int64)
    if device is not None:
        t, mn = t.to(device), mn.to(device)
    t[MIN_SLOT] = 0
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    out = t.cpu().numpy()
    out[MIN_SLOT] = 0 if int(mn.item()) == big else int(mn.item())
    return out

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


def all_reduce_vector(vec, min_slots=(), group=None, device=None) -> np.ndarray:
    """One all-reduce(SUM) over an int64 vector; the slots in ``min_slots`` hold minima where 0 means
    "nothing seen yet" and ride along as a second tiny MIN reduce."""
    import torch
    import torch.distributed as dist

    c = np.asarray(vec, dtype=np.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return c.copy()
    big = np.iinfo(np.int64).max
    t = torch.from_numpy(c.copy())
    mn = torch.tensor([c[s] if c[s] > 0 else big for s in min_slots] or [big], dtype=torch.int64)
    for s in min_slots:
        t[s] = 0
    if device is not None and dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", device) if isinstance(device, int) else device
        t, mn = t.to(dev), mn.to(dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    out = t.cpu().numpy()
    mn = mn.cpu().numpy()
    for k, s in enumerate(min_slots):
        out[s] = 0 if int(mn[k]) == big else int(mn[k])
    return out


def all_reduce_counters(counters, group=None, device=None) -> np.ndarray:
    """The path's single collective: the counter vector of ``sample_and_decode`` (slot 7 is a minimum)."""
    return all_reduce_vector(counters, min_slots=(MIN_SLOT,), group=group, device=device)


def broadcast_from_rank0(value: int, group=None, device=None) -> int:
    """Every rank returns rank 0's integer (the run's Philox seed, a stop decision...).  A no-op without an
    initialised process group."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(value)
    t = torch.tensor([int(value)], dtype=torch.int64)
    if dist.get_backend(group) == "nccl":
        t = t.to(torch.device("cuda", device) if isinstance(device, int) else (device or torch.device("cuda", torch.cuda.current_device())))
    dist.broadcast(t, src=0, group=group)
    return int(t.item())

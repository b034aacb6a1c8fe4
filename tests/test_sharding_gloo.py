"""world_size-2 gloo test of the N>1 path: shots shard by global index, one counter all-reduce.

Each rank runs the CPU oracle (test infrastructure) over its shard of the same Philox-sampled shot
set; after the collective both ranks hold the counters a single process gets for the whole set."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from bp_osd_b200.sharding import COUNTER_NAMES, MIN_SLOT, merge_counters, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL, SEED, P = 301, 1234, 0.09


def test_shard_range_partitions():
    for total in (0, 1, 7, 301, 10_000_000):
        for world in (1, 2, 3, 8):
            got = [shard_range(total, r, world) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == total
            for (s0, c0), (s1, _c1) in zip(got, got[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in got) - min(c for _, c in got) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_merge_counters_min_slot():
    a = np.array([10, 9, 9, 9, 9, 1, 50, 0]); b = np.array([5, 4, 4, 4, 4, 1, 20, 6])
    c = merge_counters(a, b)
    assert c[0] == 15 and c[MIN_SLOT] == 6
    assert merge_counters(c, np.array([1, 1, 1, 1, 1, 0, 3, 4]))[MIN_SLOT] == 4
    assert len(COUNTER_NAMES) == 8


def _counters_for(start, count):
    sys.path.insert(0, ROOT)
    from bp_osd_b200 import codes
    from oracle import oracle as O
    code = codes.config_code(1)
    n = code.N
    z = np.zeros(n)
    ex, _ = O.sample_errors(SEED, start, count, z, np.full(n, P), z)
    dec = O.OracleDecoder(code.hz, error_rate=P, max_iter=0, bp_method="ms", ms_scaling_factor=0,
                          osd_method="osd_cs", osd_order=7)
    out = dec.decode_batch(dec.syndrome(ex), want_llr=False)
    c = np.zeros(8, dtype=np.int64)
    c[0] = count
    c[1] = int(out["converge"].sum())
    fw = O.logical_fail(code.lz, ex, out["osdw"]).astype(bool)
    f0 = O.logical_fail(code.lz, ex, out["osd0"]).astype(bool)
    fb = O.logical_fail(code.lz, ex, out["bp"]).astype(bool)
    c[2] = int((out["converge"].astype(bool) & ~fb).sum())
    c[3] = int((~f0).sum())
    c[4] = int((~fw).sum())
    c[5] = int((1 - out["converge"]).sum())
    c[6] = int(out["iter"].sum())
    wts = np.concatenate([(ex ^ out["osdw"])[fw].sum(1), (ex ^ out["osd0"])[f0].sum(1)])
    c[7] = int(wts.min()) if wts.size else 0
    return c


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from bp_osd_b200.sharding import all_reduce_counters, shard_range as sr
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = sr(TOTAL, rank, world)
    total = all_reduce_counters(_counters_for(start, count))
    # the two rank-0 decisions of the harness (css_decode_sim): the Philox seed of a run that draws its own, and the
    # "save interval elapsed" flag that rides in a spare counter slot -- every rank must end up with rank 0's view
    from bp_osd_b200.sharding import all_reduce_vector, broadcast_from_rank0
    seed = broadcast_from_rank0(1000 + rank)
    flag = np.zeros(8, dtype=np.int64)
    flag[7] = 1 if rank == 0 else 0
    gate = int(all_reduce_vector(flag, min_slots=(6,))[7])
    total = np.concatenate([total, [seed, gate]])
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, total.tolist()))


def test_two_rank_counters_equal_single_process():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = _counters_for(0, TOTAL).tolist()
    assert got[0][8:] == [1000, 1] and got[1][8:] == [1000, 1]   # rank 0's seed and gate flag on both ranks
    got = {r: v[:8] for r, v in got.items()}
    assert got[0] == want and got[1] == want
    assert want[0] == TOTAL and 0 < want[4] <= TOTAL

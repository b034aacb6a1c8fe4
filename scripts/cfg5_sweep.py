"""BASELINE config 5: batch-size sweep of the large-H path (40k-qubit HGP hz, m = 19 200, n = 40 000, BP min-sum
alpha = 1 - 2^-it, max_iter = n, OSD-0) at N GPUs (strong scaling: the batch is split over the ranks).

    python scripts/cfg5_sweep.py --batches 1 8 64 512 4096 32768 262144 1048576
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/cfg5_sweep.py --batches 4096 1048576

Syndromes come from the device Philox sampler (global shot index, bit-packed); every shot is decoded through the public
`decode_batch(cuda tensor, packed=True)` call in chunks (bit-packed osdw decoding + converge + iter written for every shot);
time = CUDA events over all chunks of the rank, max over ranks.  One JSON line per batch size on rank 0.
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bp_osd_b200 import codes, BpOsdDecoder
from bp_osd_b200.sharding import shard_range

ap = argparse.ArgumentParser()
ap.add_argument("--batches", type=int, nargs="*", default=[1, 8, 64, 512, 4096, 32768, 262144, 1048576])
ap.add_argument("--p", type=float, default=0.02)
ap.add_argument("--chunk", type=int, default=49152)
ap.add_argument("--precision", type=int, default=64)
ap.add_argument("--cfg", type=int, default=5)
a = ap.parse_args()
sys.stdout.flush()
REAL_STDOUT = os.dup(1)   # JSON lines go here; whatever libraries print (NCCL banner) goes to stderr
os.dup2(2, 1)
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
H = codes.config_code(a.cfg, logicals=False).hz
m, n = H.shape
dec = BpOsdDecoder(H, error_rate=a.p, max_iter=0, bp_method="ms", ms_scaling_factor=0, osd_method="osd0", osd_order=0,
                   precision=a.precision, device=local)
dec.set_error_channel(px=a.p)
info = dec.info()
nb = (n + 7) // 8
out = {"osdw": torch.empty((a.chunk, nb), dtype=torch.uint8, device=dev), "converge": torch.empty(a.chunk, dtype=torch.uint8, device=dev),
       "iter": torch.empty(a.chunk, dtype=torch.int32, device=dev)}


def run(B, shot0):
    lo, cnt = shard_range(B, rank, world)   # (start, count) of this rank's share of the batch
    hi = lo + cnt
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = conv = osd = 0
    ms_bp = ms_osd = 0.0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for c0 in range(lo, hi, a.chunk):
        Bc = min(a.chunk, hi - c0)
        _, syn = dec.sample_syndromes(0xB905D, shot0 + c0, Bc, sector=0, return_errors=False, packed=True)
        o = {k: v[:Bc] for k, v in out.items()}
        dec.decode_batch(syn, return_llr=False, return_all=False, out=o, packed=True)
        st = dec.stats()
        iters += st["bp_iterations"]; conv += st["bp_converged"]; osd += st["osd_invocations"]
        ms_bp += st["ms_bp"]; ms_osd += st["ms_osd"]
    ev1.record()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1), ms_bp, ms_osd], dtype=torch.float64, device=dev)
    c = torch.tensor([iters, conv, osd, hi - lo], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()], [int(x) for x in c.tolist()]


run(min(a.batches[0] if a.batches else 64, 256) or 1, 0)  # warm-up (workspaces, first launches)
shot0 = 10**7
for B in a.batches:
    (ms, ms_bp, ms_osd), (iters, conv, osd, shots) = run(B, shot0)
    shot0 += B
    if rank == 0:
        os.write(REAL_STDOUT, (json.dumps({"config": 5 if a.cfg == 5 else a.cfg, "n_gpus": world, "batch": B, "ms": ms, "shots_per_s": B / (ms * 1e-3),
                          "bp_ms_max_rank": ms_bp, "osd_ms_max_rank": ms_osd, "mean_iterations": iters / max(shots, 1),
                          "bp_converged_frac": conv / max(shots, 1), "osd_shots": osd,
                          "bp_shot_iterations_per_s": iters / (ms_bp * 1e-3) if ms_bp > 0 else None,
                          "p": a.p, "precision": a.precision, "bp_kernel": info["bp_kernel"], "bp_cluster_size": info["bp_cluster_size"],
                          "osd_variant": info["osd_variant"]}) + "\n").encode())
if world > 1:
    dist.destroy_process_group()

#!/usr/bin/env python
"""bench.py -- shots/sec of BP+OSD-CS(7) on the [[1922,50,16]] hypergraph-product code (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this framework, N ranks (torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle restatement of ldpc
                                                             # (ldpc itself is not installable offline)

One step = one pass of the decode hot path over one batch of synthetic syndromes per GPU
(weak scaling: the per-GPU batch is fixed; at 8 GPUs the job is BASELINE's 10M shots).  Syndromes
come from the device Philox sampler with global shot indices, so the union over ranks is the same
shot set for any N.  `value` is timed with the syndromes resident in HBM; `e2e` goes through the
public `decode_batch(numpy)` call with pinned host buffers, H2D and D2H inside the timed region.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(cfg=3, p=0.05, max_iter=0, bp_method="ms", ms_scaling_factor=0, osd_method="osd_cs", osd_order=7)
SEED = 0xB905D


def workload_name(shots, prec):
    return (f"[[1922,50,16]] HGP hz sector (m=961,n=1922,E=5766), bit-flip p=0.05, BP min-sum alpha=1-2^-it "
            f"max_iter=n + OSD-CS order 7, {shots} shots/GPU/step, fp{prec}")


def algorithmic_bp_bytes(n, m, E, iterations, shots, w):
    """it*(4E+2n)*w per shot + (ceil(m/8) + n + n*w) in/out."""
    return iterations * (4 * E + 2 * n) * w + shots * ((m + 7) // 8 + n + n * w)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle on all host cores
# ------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init():
    from bp_osd_b200 import codes
    from oracle.oracle import OracleDecoder
    H = codes.config_code(CFG["cfg"], logicals=False).hz if CFG["cfg"] != 3 else codes.config_code(3).hz
    _W["dec"] = OracleDecoder(H, error_rate=CFG["p"], max_iter=CFG["max_iter"], bp_method=CFG["bp_method"],
                              ms_scaling_factor=CFG["ms_scaling_factor"], osd_method=CFG["osd_method"],
                              osd_order=CFG["osd_order"])
    _W["n"] = H.shape[1]


def _cpu_work(args):
    from oracle.oracle import sample_errors
    shot0, shots = args
    dec = _W["dec"]
    z = np.zeros(_W["n"])
    ex, _ = sample_errors(SEED, shot0, shots, z, np.full(_W["n"], CFG["p"]), z)
    s = dec.syndrome(ex)
    t = time.perf_counter()
    x0, n0 = dec.totals
    out = dec.decode_batch(s, want_llr=False)
    dt = time.perf_counter() - t
    x1, n1 = dec.totals
    return shots, dt, int(out["converge"].sum()), x1 - x0, n1 - n0


class CpuArm:
    def __init__(self):
        import multiprocessing as mp
        self.cores = len(os.sched_getaffinity(0))
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=_cpu_init)

    def step(self, shot0, shots_per_core):
        jobs = [(shot0 + c * shots_per_core, shots_per_core) for c in range(self.cores)]
        t = time.perf_counter()
        res = self.pool.map(_cpu_work, jobs)
        wall = time.perf_counter() - t
        self.elim_wordxors = sum(r[3] for r in res)   # algorithmic elimination ops counted by the oracle
        self.osd_shots = sum(r[4] for r in res)
        return sum(r[0] for r in res), wall

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as _o
    _o.build()
    arm = CpuArm()
    per_core = args.cpu_shots_per_core
    shot0 = 0
    for _ in range(args.warmup):
        n, _w = arm.step(shot0, max(per_core // 4, 50)); shot0 += n
    tot, wall = 0, 0.0
    for _ in range(args.steps):
        n, w = arm.step(shot0, per_core); shot0 += n
        tot += n; wall += w
    arm.close()
    v = tot / wall
    sample = f"{per_core} shots/core/step x {arm.cores} cores x {args.steps} steps of the same workload (Philox seed {SEED:#x})"
    line = {
        "impl": "reference", "metric": "shots/sec BP+OSD-CS(7) on [[1922,50,16]] HGP", "value": v, "unit": "shots/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(per_core * arm.cores, 64),
                   "note": "CPU restatement of ldpc v2 bposd_decoder (oracle/bposd_oracle.c); ldpc is not installable offline"},
        "cpu_baseline": {"value": v, "unit": "shots/s", "cores": arm.cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from bp_osd_b200 import codes, BpOsdDecoder

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the decoder has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own messages (its version banner, NCCL_DEBUG output) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    code = codes.config_code(CFG["cfg"])
    H = code.hz
    m, n = H.shape
    E = H.nnz
    prec = args.precision
    w = prec // 8
    dec = BpOsdDecoder(H, error_rate=CFG["p"], max_iter=CFG["max_iter"], bp_method=CFG["bp_method"],
                       ms_scaling_factor=CFG["ms_scaling_factor"], osd_method=CFG["osd_method"],
                       osd_order=CFG["osd_order"], precision=prec, device=local)
    if args.bp_kernel is not None or args.bp_threads:
        dec.set_tuning(bp_kernel=args.bp_kernel, bp_threads=args.bp_threads)
    dec.set_error_channel(px=CFG["p"])
    dec.set_logicals(code.lz)
    info = dec.info()
    S = args.shots_per_gpu
    total_steps = args.warmup + args.steps

    # inputs: one fresh batch of syndromes per step, sampled on the device BEFORE the timed region and
    # kept resident in HBM; each batch (S x 961 B, 1.2 GB at the default S) is larger than the 126 MB L2.
    syn_batches, err_last = [], None
    for step in range(total_steps):
        shot0 = (step * world + rank) * S
        want_err = step == total_steps - 1
        e_, s_ = dec.sample_syndromes(SEED, shot0, S, sector=0, return_errors=want_err)
        syn_batches.append(s_)
        if want_err:
            err_last = e_
    tdt = torch.float64 if prec == 64 else torch.float32
    bufs = {"osdw": torch.empty((S, n), dtype=torch.uint8, device=dev), "osd0": torch.empty((S, n), dtype=torch.uint8, device=dev),
            "bp": torch.empty((S, n), dtype=torch.uint8, device=dev), "llr": torch.empty((S, n), dtype=tdt, device=dev),
            "converge": torch.empty(S, dtype=torch.uint8, device=dev), "iter": torch.empty(S, dtype=torch.int32, device=dev)}
    torch.cuda.synchronize()

    # ---- device-resident throughput ----
    clocks = ClockSampler(local)
    ms_bp = ms_osd = 0.0
    iters = conv = osd_inv = launches = 0
    counters = torch.zeros(8, dtype=torch.int64, device=dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    res = None
    t_wall0 = time.time()
    for step in range(total_steps):
        if step == args.warmup:
            barrier()
            clocks.start()
            t_wall0 = time.time()
            ev0.record()
        res = dec.decode_batch(syn_batches[step], return_llr=True, return_all=True, out=bufs)
        if step >= args.warmup:
            st = dec.stats()
            ms_bp += st["ms_bp"]; ms_osd += st["ms_osd"]
            iters += st["bp_iterations"]; conv += st["bp_converged"]; osd_inv += st["osd_invocations"]
            launches += st["launches"]
    # the path's single collective: the logical-failure counters of the last step
    fail = dec.logical_check(err_last, res.osdw_decoding)
    counters[0] = S
    counters[1] = int(fail.sum())
    launches += 1
    if world > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    ms_total = ev0.elapsed_time(ev1)
    tt = torch.tensor([ms_total, ms_bp, ms_osd], dtype=torch.float64, device=dev)
    agg = torch.tensor([iters, conv, osd_inv, launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    ms_total, ms_bp_max, ms_osd_max = (float(x) for x in tt.tolist())
    iters_all, conv_all, osd_all, launches_all = (int(x) for x in agg.tolist())
    clk = clocks.stop(t_wall0, t_wall1)
    shots_all = S * world * args.steps
    value = shots_all / (ms_total * 1e-3)
    del res, err_last, syn_batches, bufs
    torch.cuda.empty_cache()

    # ---- roofline of the dominant kernel (BP), from CUDA events on the launching stream ----
    peak, peak_src = measured_peak()
    alg_bytes_rank = algorithmic_bp_bytes(n, m, E, iters, S * args.steps, w)
    achieved = alg_bytes_rank / (ms_bp * 1e-3) / 1e9 if ms_bp > 0 else None
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tj = json.load(f)
        per_shot = tj.get(f"bp_fp{prec}_dram_bytes_per_shot")
        traffic = per_shot * S if per_shot else None  # ncu --set full capture, scaled to this launch's shots
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": f"bp kernel variant {info['bp_kernel']} (fp{prec})", "achieved": achieved, "peak": peak,
        "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes_rank / max(args.steps, 1),
        "note": "algorithmic bytes = it*(4E+2n)*w + in/out per SURVEY 8(d); messages are kept in shared memory, so "
                "frac can exceed 1: HBM is not the binding resource, shared-memory bandwidth is (see DESIGN.md)",
        "smem": {"bound": "shared-memory pipe", "achieved": (iters * 4 * E * w) / (ms_bp * 1e-3) / 1e9 if ms_bp > 0 else None,
                 "peak": info["sm_count"] * 128 * (clk.get("sm_max_mhz") or 1965.0) * 1e6 / 1e9, "unit": "GB/s",
                 "note": "4*E*w bytes per shot-iteration (each message read and written once per sweep) against "
                         "SMs x 128 B/clk x max SM clock"},
        "bp_ms_per_step": ms_bp / args.steps, "osd_ms_per_step": ms_osd / args.steps,
        "mean_iterations": iters / (S * args.steps), "bp_shot_iterations_per_s": iters / (ms_bp * 1e-3) if ms_bp else None,
    }

    if roofline["smem"]["achieved"]:
        roofline["smem"]["frac"] = roofline["smem"]["achieved"] / roofline["smem"]["peak"]

    # ---- end to end through the public API with pinned host buffers ----
    Se = min(S, args.e2e_shots_per_gpu)
    h_syn = torch.empty((Se, m), dtype=torch.uint8, pin_memory=True)
    h_out = {"osdw": torch.empty((Se, n), dtype=torch.uint8, pin_memory=True).numpy(),
             "converge": torch.empty(Se, dtype=torch.uint8, pin_memory=True).numpy(),
             "iter": torch.empty(Se, dtype=torch.int32, pin_memory=True).numpy()}
    e2e_ms = 0.0
    for step in range(total_steps):
        _, syn = dec.sample_syndromes(SEED, (step * world + rank) * Se, Se, sector=0, return_errors=False)
        h_syn.copy_(syn)
        torch.cuda.synchronize()
        del syn
        if step == args.warmup:
            barrier()
            t0 = time.perf_counter()
        dec.decode_batch(h_syn.numpy(), return_llr=False, return_all=False, out=h_out)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t2 = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = Se * world * args.steps / (float(t2.item()) * 1e-3)
    e2e = {"value": e2e_value, "unit": "shots/s", "h2d_bytes_per_step": int(Se * m),
           "d2h_bytes_per_step": int(Se * n + Se + 4 * Se), "shots_per_gpu_per_step": Se,
           "api": "BpOsdDecoder.decode_batch(numpy uint8[B,m]) -> bposd_decode_host (C ABI), pinned host buffers"}

    # ---- single-shot latency of decode() (B=1: launch + D2H) ----
    lat = None
    cpu_baseline = None
    if rank == 0:
        _, syn1 = dec.sample_syndromes(SEED, 10**9, 300, sector=0, return_errors=False)
        s1 = syn1.cpu().numpy()
        ts = []
        for i in range(300):
            t = time.perf_counter(); dec.decode(s1[i]); ts.append(time.perf_counter() - t)
        ts = np.array(ts[50:]) * 1e6
        lat = {"p50_us": float(np.percentile(ts, 50)), "p99_us": float(np.percentile(ts, 99)), "samples": int(ts.size),
               "path": "decoder.decode(syndrome) -> bposd_decode_host B=1: one kernel launch + one stream synchronise, syndrome "
                       "read from and results written to pinned host memory by the kernel, one SM per shot (latency geometry)"}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as _o
            _o.build()
            arm = CpuArm()
            arm.step(0, 50)
            nshots, wall = arm.step(10**6, args.cpu_shots_per_core)
            arm.close()
            if arm.osd_shots:
                # OSD roofline: algorithmic 32-bit word ops per OSD shot = bit-packed elimination word
                # XORs as counted by the oracle on this sample + candidates x (XOR + POPC) x ceil(rank/32)
                ncand = info["k"] + CFG["osd_order"] * (CFG["osd_order"] - 1) // 2
                ops_shot = arm.elim_wordxors / arm.osd_shots + ncand * 2 * ((info["rank"] + 31) // 32)
                peak_ops = dec.int32_peak()
                ach = ops_shot * (osd_inv / args.steps) / (ms_osd / args.steps * 1e-3) if ms_osd > 0 else None
                roofline["osd"] = {"bound": "int32 alu (LOP3)", "achieved": ach, "peak": peak_ops, "unit": "ops/s",
                                   "frac": ach / peak_ops if ach else None,
                                   "algorithmic_ops_per_osd_shot": ops_shot, "osd_shots_sampled": arm.osd_shots,
                                   "note": "peak = LOP3 microbenchmark in this run; ops per shot from the oracle's count"}
            cpu_baseline = {"value": nshots / wall, "unit": "shots/s", "cores": arm.cores, "kind": "port",
                            "sample": f"{args.cpu_shots_per_core} shots/core x {arm.cores} cores of the same workload, "
                                      "oracle/bposd_oracle.c (restatement of ldpc v2; ldpc not installable offline)"}

    if rank == 0:
        line = {
            "metric": "shots/sec BP+OSD-CS(7) on [[1922,50,16]] HGP", "value": value, "unit": "shots/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": f"f{prec}", "data": "synthetic",
            "config": {"workload": workload_name(S, prec), "global_shots_per_step": S * world,
                       "l2": "each step decodes a fresh syndrome batch larger than L2 (S*961 B)",
                       "outputs": "osdw, osd0, bp, llr, converge, iter written to HBM for every shot",
                       "bp_kernel": info["bp_kernel"], "bp_threads": info["bp_threads"],
                       "bp_ctas_per_sm": info["bp_ctas_per_sm"], "sm_count": info["sm_count"]},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches_all,
            "clocks": clk, "latency": lat,
            "decode_stats": {"bp_converged_frac": conv_all / shots_all, "osd_invocation_frac": osd_all / shots_all,
                             "logical_failures_last_step": int(counters[1].item()), "shots_last_step": int(counters[0].item())},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32])
    ap.add_argument("--shots-per-gpu", type=int, default=1_250_000)
    ap.add_argument("--e2e-shots-per-gpu", type=int, default=1_250_000)
    ap.add_argument("--cpu-shots-per-core", type=int, default=1000)
    ap.add_argument("--bp-kernel", type=int, default=None)
    ap.add_argument("--bp-threads", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

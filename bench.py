#!/usr/bin/env python
"""bench.py -- shots/sec of BP+OSD on the codes of BASELINE.json; default: BP+OSD-CS(7) on the [[1922,50,16]] HGP code.

    python bench.py --gpus N --steps K --warmup W            # this framework, N ranks (torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: ldpc's BpOsdDecoder when it can be imported,
                                                             # else the oracle restatement of it (ldpc is not installable offline)
    python bench.py --config 3 --ms-scaling-factor 0.625     # the harness-default scaling: half the shots reach OSD
    python bench.py --config {1,2,4,5}                       # the other BASELINE configs (see CONFIGS below)

One step = one pass of the decode hot path over one batch of synthetic syndromes per GPU (weak scaling: the per-GPU
batch is fixed; at 8 GPUs the default job is BASELINE's 10M shots).  Syndromes come from the device Philox sampler with
global shot indices, so the union over ranks is the same shot set for any N.
`value`  : syndromes resident in HBM, every output written to HBM (CUDA events, max over ranks).
`e2e`    : the public `decode_batch(host array, packed=True)` call on pre-staged pinned host batches -- H2D of the
           bit-packed syndromes and D2H of the bit-packed decodings, converge flags and iteration counts inside the
           timed region, nothing else.
`roofline`: the BP kernel against the MEASURED shared-memory bandwidth (its messages live in shared memory); the
           SURVEY 8(d) HBM figure and the measured DRAM traffic are kept beside it.
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

SEED = 0xB905D
# BASELINE.json configs[cfg-1]; decoder arguments per SURVEY.md 8(d) "Synthetic inputs"
CONFIGS = {
    1: dict(name="d=5 surface code hgp(rep_code(5)) hz [[41,1,5]]", p=0.05, max_iter=0, bp_method="ms", ms_scaling_factor=0.0,
            osd_method="osd_cs", osd_order=7, shots=1_000_000, cpu_shots=20000),
    2: dict(name="[[400,16,6]] HGP hz", p=0.05, max_iter=0, bp_method="ms", ms_scaling_factor=0.0, osd_method="osd_cs",
            osd_order=7, shots=1_000_000, cpu_shots=4000),
    3: dict(name="[[1922,50,16]] HGP hz", p=0.05, max_iter=0, bp_method="ms", ms_scaling_factor=0.0, osd_method="osd_cs",
            osd_order=7, shots=1_250_000, cpu_shots=1000),
    4: dict(name="[[882,24]] lifted-product hz", p=0.05, max_iter=0, bp_method="ps", ms_scaling_factor=0.0, osd_method="osd_e",
            osd_order=10, shots=200_000, cpu_shots=300),
    5: dict(name="40k-qubit HGP hz (m=19200, n=40000)", p=0.02, max_iter=0, bp_method="ms", ms_scaling_factor=0.0,
            osd_method="osd0", osd_order=0, shots=4096, cpu_shots=4),
}
METRIC = "shots/sec BP+OSD-CS(7) on [[1922,50,16]] HGP"


def decoder_kwargs(c):
    return dict(max_iter=c["max_iter"], bp_method=c["bp_method"], ms_scaling_factor=c["ms_scaling_factor"],
                osd_method=c["osd_method"], osd_order=c["osd_order"])


def workload_name(cfg, c, m, n, E, prec):
    """The same string in both arms (the driver compares `config` across them): no batch sizes in here."""
    alpha = "alpha=1-2^-it" if c["ms_scaling_factor"] == 0 else f"alpha={c['ms_scaling_factor']:g}"
    bp = f"BP min-sum {alpha}" if c["bp_method"] == "ms" else "BP product-sum"
    osd = {"osd_cs": f"OSD-CS order {c['osd_order']}", "osd_e": f"OSD-E order {c['osd_order']}", "osd0": "OSD-0"}[c["osd_method"]]
    return (f"config {cfg}: {c['name']} (m={m},n={n},E={E}), bit-flip p={c['p']:g}, {bp} max_iter=n + {osd}, fp{prec}")


def metric_name(cfg, c):
    if cfg == 3 and c["osd_method"] == "osd_cs" and c["osd_order"] == 7:
        return METRIC
    return f"shots/sec BP+{c['osd_method']}({c['osd_order']}) on BASELINE config {cfg}"


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


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU arm: ldpc.BpOsdDecoder when importable (baseline/_ref or site-packages), else the oracle, on all host cores
# ------------------------------------------------------------------------------------------------
_W = {}


def try_import_ldpc():
    """The reference's decoder is `from ldpc import BpOsdDecoder` (/root/reference/src/bposd/css_decode_sim.py:6)."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref) and ref not in sys.path:
        sys.path.insert(0, ref)
    try:
        import ldpc  # noqa: F401
        from ldpc import BpOsdDecoder  # noqa: F401
        return True
    except Exception:
        return False


def _cpu_init(cfg, c, use_ldpc):
    from bp_osd_b200 import codes
    H = codes.config_code(cfg, logicals=False).hz
    kw = decoder_kwargs(c)
    if use_ldpc:
        from ldpc import BpOsdDecoder as Ref
        # the constructor call of the reference harness, css_decode_sim.py:444-452
        _W["dec"] = Ref(H, channel_probs=np.full(H.shape[1], c["p"]), max_iter=kw["max_iter"] or H.shape[1],
                        bp_method=kw["bp_method"], ms_scaling_factor=float(kw["ms_scaling_factor"]), osd_method=kw["osd_method"],
                        osd_order=kw["osd_order"])
    else:
        from oracle.oracle import OracleDecoder
        _W["dec"] = OracleDecoder(H, error_rate=c["p"], math="libm", **kw)  # libm: the arithmetic ldpc itself runs
    from oracle.oracle import OracleDecoder as _O
    _W["syn"] = _O(H, error_rate=c["p"], osd_method="osd0")
    _W["n"], _W["p"], _W["ldpc"] = H.shape[1], c["p"], use_ldpc


def _cpu_work(args):
    from oracle.oracle import sample_errors
    shot0, shots = args
    dec = _W["dec"]
    z = np.zeros(_W["n"])
    ex, _ = sample_errors(SEED, shot0, shots, z, np.full(_W["n"], _W["p"]), z)
    s = _W["syn"].syndrome(ex)
    if _W["ldpc"]:
        t = time.perf_counter()
        for b in range(shots):
            dec.decode(s[b])          # the per-shot call of css_decode_sim.py:174
        return shots, time.perf_counter() - t, 0, 0, 0
    x0, n0 = dec.totals
    t = time.perf_counter()          # sampling and syndromes above are not part of the decode path that is timed
    out = dec.decode_batch(s, want_llr=False)
    dt = time.perf_counter() - t
    x1, n1 = dec.totals
    return shots, dt, int(out["converge"].sum()), x1 - x0, n1 - n0


class CpuArm:
    def __init__(self, cfg, c):
        import multiprocessing as mp
        from functools import partial
        self.cores = len(os.sched_getaffinity(0))
        self.ldpc = try_import_ldpc()
        self.kind = "ldpc" if self.ldpc else "port"
        self.what = ("ldpc.BpOsdDecoder.decode per shot (the reference's own decoder)" if self.ldpc else
                     "oracle/bposd_oracle.c -O3 (restatement of ldpc v2; ldpc is not installable offline)")
        self.pool = mp.get_context("fork").Pool(self.cores, initializer=partial(_cpu_init, cfg, c, self.ldpc))

    def step(self, shot0, shots_per_core):
        jobs = [(shot0 + k * shots_per_core, shots_per_core) for k in range(self.cores)]
        res = self.pool.map(_cpu_work, jobs)
        self.elim_wordxors = sum(r[3] for r in res)   # algorithmic elimination ops counted by the oracle
        self.osd_shots = sum(r[4] for r in res)
        # all cores decode concurrently: the step lasts as long as its slowest worker's decode loop
        return sum(r[0] for r in res), max(r[1] for r in res)

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args, cfg, c):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bp_osd_b200 import codes
    from oracle import oracle as _o
    _o.build()
    H = codes.config_code(cfg, logicals=False).hz
    m, n = H.shape
    arm = CpuArm(cfg, c)
    per_core = args.cpu_shots_per_core or c["cpu_shots"]
    shot0 = 0
    for _ in range(args.warmup):
        k, _w = arm.step(shot0, max(per_core // 4, 1)); shot0 += k
    tot, wall = 0, 0.0
    for _ in range(args.steps):
        k, w = arm.step(shot0, per_core); shot0 += k
        tot += k; wall += w
    arm.close()
    v = tot / wall
    sample = f"{per_core} shots/core/step x {arm.cores} cores x {args.steps} steps of the same workload (Philox seed {SEED:#x}); {arm.what}"
    line = {
        "impl": "reference", "metric": metric_name(cfg, c), "value": v, "unit": "shots/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(cfg, c, m, n, H.nnz, 64)},
        "shots_per_step": per_core * arm.cores,
        "cpu_baseline": {"value": v, "unit": "shots/s", "cores": arm.cores, "kind": arm.kind, "sample": sample},
        "e2e": {"value": v, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args, cfg, c):
    # stdout carries exactly one JSON line: whatever libraries print while the job runs (NCCL's version banner, warnings of
    # the process-group backend) is sent to stderr at the file-descriptor level; the line is written to the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
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

    code = codes.config_code(cfg, logicals=cfg != 5)
    H = code.hz
    m, n = H.shape
    E = H.nnz
    prec = args.precision
    w = prec // 8
    dec = BpOsdDecoder(H, error_rate=c["p"], precision=prec, device=local, **decoder_kwargs(c))
    if args.bp_kernel is not None or args.bp_threads:
        dec.set_tuning(bp_kernel=args.bp_kernel, bp_threads=args.bp_threads)
    if args.osd_variant is not None:
        dec.set_osd_variant(args.osd_variant)
    dec.set_error_channel(px=c["p"])
    have_logicals = getattr(code, "lz", None) is not None and cfg != 5
    if have_logicals:
        dec.set_logicals(code.lz)
    info = dec.info()
    S = args.shots_per_gpu or c["shots"]
    total_steps = args.warmup + args.steps

    # inputs: one fresh batch of syndromes per step, sampled on the device BEFORE the timed region and
    # kept resident in HBM; each batch (S x m bytes, 1.2 GB at the default) is larger than the 126 MB L2.
    syn_batches, err_last = [], None
    for step in range(total_steps):
        shot0 = (step * world + rank) * S
        want_err = step == total_steps - 1 and have_logicals
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
        if step == 0 and args.warmup > 0 and have_logicals:
            # warm-up of the end-of-run logical check too (first launch of its kernel and of the reduction: module loads
            # that otherwise land inside the timed region, once per run); the pairing with err_last is immaterial here
            int(dec.logical_check(err_last, res.osdw_decoding).sum())
        if step >= args.warmup:
            st = dec.stats()
            ms_bp += st["ms_bp"]; ms_osd += st["ms_osd"]
            iters += st["bp_iterations"]; conv += st["bp_converged"]; osd_inv += st["osd_invocations"]
            launches += st["launches"]
    # the path's single collective: the logical-failure counters of the last step
    counters[0] = S
    if have_logicals:
        fail = dec.logical_check(err_last, res.osdw_decoding)
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
    # Bound: shared-memory bandwidth.  Every edge message is read and written once per sweep, 4*E*w bytes per
    # shot-iteration, and the messages never leave shared memory (or the cluster's distributed shared memory, kernel 3);
    # the peak is measured in this run (conflict-free 16-byte LDS + STS, the kernel's own access mix).
    smem_peak = dec.smem_peak() / 1e9
    smem_bytes = iters * 4 * E * w
    achieved = smem_bytes / (ms_bp * 1e-3) / 1e9 if ms_bp > 0 else None
    hbm_peak, hbm_src = measured_hbm_peak()
    alg_bytes_rank = algorithmic_bp_bytes(n, m, E, iters, S * args.steps, w)
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            tj = json.load(f)
        per_shot = tj.get(f"bp_fp{prec}_dram_bytes_per_shot") if cfg == 3 else None
        traffic = per_shot * S if per_shot else None  # ncu --set full capture, scaled to this launch's shots
    except Exception:
        pass
    hbm_ach = alg_bytes_rank / (ms_bp * 1e-3) / 1e9 if ms_bp > 0 else None
    roofline = {
        "bound": "smem", "kernel": f"bp kernel variant {info['bp_kernel']} (fp{prec})", "achieved": achieved, "peak": smem_peak,
        "unit": "GB/s", "frac": (achieved / smem_peak) if achieved else None, "traffic": traffic,
        "peak_source": "measured in this run: bposd_smem_peak (16-byte LDS+STS, conflict free, all SMs)",
        "peak_theoretical": info["sm_count"] * 128 * (clk.get("sm_max_mhz") or 1965.0) * 1e6 / 1e9,
        "algorithmic_bytes_per_launch": smem_bytes / max(args.steps, 1),
        "note": "achieved = 4*E*w bytes per shot-iteration (each message read and written once per sweep) / CUDA-event time of "
                "the BP launches; traffic = measured DRAM bytes per launch (ncu), i.e. only inputs and results touch HBM",
        "hbm": {"bound": "hbm (SURVEY 8d formula; not the binding resource)", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": (hbm_ach / hbm_peak) if hbm_ach else None, "peak_source": hbm_src,
                "algorithmic_bytes_per_launch": alg_bytes_rank / max(args.steps, 1),
                "note": "it*(4E+2n)*w + in/out bytes per SURVEY 8(d) against the measured copy bandwidth; above 1 because the "
                        "messages are staged in shared memory"},
        "bp_ms_per_step": ms_bp / args.steps, "osd_ms_per_step": ms_osd / args.steps,
        "mean_iterations": iters / (S * args.steps), "bp_shot_iterations_per_s": iters / (ms_bp * 1e-3) if ms_bp else None,
    }

    if c["bp_method"] == "ps" and prec == 64:
        # Product-sum is not memory bound: tanh, log and three divisions per edge make it an fp64-pipe kernel.  fp64
        # operations per edge and iteration, counted from include/bposd_math.h and the update's own arithmetic: b2c / 2 (1),
        # tanh (37: expm1 25, u + 2, eight for the division sequence, 1 - q, NaN select), forward / reverse products (3),
        # (1 + x) / (1 - x) (11), sign (1), two additions on the bit side = 55 on EVERY edge, plus log (32) = 87 -- except
        # that log leaves through its special-value exit when the product of tanh has saturated to +-1 (argument 0 or inf),
        # which is the normal state of the shots that do not converge, and those run max_iter passes and so carry most
        # of the iterations of this workload.  `achieved` therefore uses the 55 operations every edge executes (a lower
        # bound); `frac_full_path` is the same with 87 (an upper bound).  The hardware counter of the same kernel
        # (sm__pipe_fp64_cycles_active, profiles/r2ac_ps_ncu_full_summary.json) reads 0.66.  Peak: DFMA issue rate measured
        # in this run.
        PS_FP64_OPS_MIN, PS_FP64_OPS_FULL = 55, 87
        fp64_peak = dec.fp64_peak() / 1e12
        ops = iters * E * PS_FP64_OPS_MIN
        ach = ops / (ms_bp * 1e-3) / 1e12 if ms_bp > 0 else None
        roofline = {"bound": "fp64", "kernel": roofline["kernel"], "achieved": ach, "peak": fp64_peak, "unit": "T fp64 op/s",
                    "frac": (ach / fp64_peak) if ach else None,
                    "frac_full_path": (ach / fp64_peak * PS_FP64_OPS_FULL / PS_FP64_OPS_MIN) if ach else None, "traffic": traffic,
                    "peak_source": "measured in this run: bposd_fp64_peak (independent DFMA chains, all SMs)",
                    "algorithmic_ops_per_launch": ops / max(args.steps, 1),
                    "note": f"achieved = {PS_FP64_OPS_MIN} fp64 operations per edge and iteration (executed on every edge; {PS_FP64_OPS_FULL} when "
                            "log takes its main path, see bench.py) x E x shot-iterations / CUDA-event time of the BP launches",
                    "smem": {k: roofline[k] for k in ("bound", "achieved", "peak", "unit", "frac", "peak_source")},
                    "hbm": roofline["hbm"], "bp_ms_per_step": roofline["bp_ms_per_step"], "osd_ms_per_step": roofline["osd_ms_per_step"],
                    "mean_iterations": roofline["mean_iterations"], "bp_shot_iterations_per_s": roofline["bp_shot_iterations_per_s"]}

    # ---- end to end through the public API: pre-staged pinned host batches, bit-packed both ways ----
    Se = min(S, args.e2e_shots_per_gpu or S)
    mb, nb = (m + 7) // 8, (n + 7) // 8
    nstage = min(total_steps, 3)
    h_syn = []
    for k in range(nstage):   # staged BEFORE the timed region; the timed loop rotates over them
        _, syn = dec.sample_syndromes(SEED, ((total_steps + k) * world + rank) * Se, Se, sector=0, return_errors=False, packed=True)
        hs = torch.empty((Se, mb), dtype=torch.uint8, pin_memory=True)
        hs.copy_(syn)
        h_syn.append(hs.numpy())
        del syn
    torch.cuda.synchronize()
    h_out = {"osdw": torch.empty((Se, nb), dtype=torch.uint8, pin_memory=True).numpy(),
             "converge": torch.empty(Se, dtype=torch.uint8, pin_memory=True).numpy(),
             "iter": torch.empty(Se, dtype=torch.int32, pin_memory=True).numpy()}
    t0 = time.perf_counter()
    for step in range(total_steps):
        if step == args.warmup:
            barrier()
            t0 = time.perf_counter()
        dec.decode_batch(h_syn[step % nstage], return_llr=False, return_all=False, out=h_out, packed=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t2 = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = Se * world * args.steps / (float(t2.item()) * 1e-3)
    e2e = {"value": e2e_value, "unit": "shots/s", "h2d_bytes_per_step": int(Se * mb),
           "d2h_bytes_per_step": int(Se * nb + Se + 4 * Se), "shots_per_gpu_per_step": Se,
           "api": "BpOsdDecoder.decode_batch(numpy uint8[B, ceil(m/8)], packed=True) -> bposd_decode_host_packed (C ABI); "
                  "pinned host buffers staged before the timed region; results: packed osdw decoding, converge, iter"}
    del h_syn, h_out

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
               "path": "decoder.decode(syndrome) -> bposd_decode_host B=1"}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as _o
            _o.build()
            arm = CpuArm(cfg, c)
            per_core = args.cpu_shots_per_core or c["cpu_shots"]
            arm.step(0, max(per_core // 20, 1))
            nshots, wall = arm.step(10**6, per_core)
            arm.close()
            if arm.osd_shots and osd_inv and ms_osd > 0:
                # OSD roofline: algorithmic 32-bit word ops per OSD shot = bit-packed elimination word
                # XORs as counted by the oracle on this sample + candidates x (XOR + POPC) x ceil(rank/32)
                ncand = {"osd_cs": info["k"] + c["osd_order"] * (c["osd_order"] - 1) // 2, "osd_e": 2 ** c["osd_order"] - 1,
                         "osd0": 0}[c["osd_method"]]
                ops_shot = arm.elim_wordxors / arm.osd_shots + ncand * 2 * ((info["rank"] + 31) // 32)
                peak_ops = dec.int32_peak()
                ach = ops_shot * (osd_inv / args.steps) / (ms_osd / args.steps * 1e-3)
                roofline["osd"] = {"bound": "int32 alu (LOP3)", "kernel": f"osd kernel variant {info['osd_variant']}", "achieved": ach,
                                   "peak": peak_ops, "unit": "ops/s", "frac": ach / peak_ops,
                                   "algorithmic_ops_per_osd_shot": ops_shot, "osd_shots_sampled": arm.osd_shots,
                                   "osd_shots_per_s": (osd_inv / args.steps) / (ms_osd / args.steps * 1e-3),
                                   "note": "peak = LOP3 microbenchmark in this run; ops per shot from the oracle's count"}
            cpu_baseline = {"value": nshots / wall, "unit": "shots/s", "cores": arm.cores, "kind": arm.kind,
                            "sample": f"{per_core} shots/core x {arm.cores} cores of the same workload; {arm.what}"}

    if rank == 0:
        line = {
            "metric": metric_name(cfg, c), "value": value, "unit": "shots/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": f"f{prec}", "data": "synthetic",
            "config": {"workload": workload_name(cfg, c, m, n, E, prec)},
            "shots_per_step": S * world,
            "run": {"shots_per_gpu_per_step": S, "l2": "each step decodes a fresh syndrome batch larger than L2 (S*m bytes)",
                    "outputs": "osdw, osd0, bp, llr, converge, iter written to HBM for every shot",
                    "bp_kernel": info["bp_kernel"], "bp_threads": info["bp_threads"], "bp_ctas_per_sm": info["bp_ctas_per_sm"],
                    "osd_variant": info["osd_variant"], "sm_count": info["sm_count"]},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches_all,
            "clocks": clk, "latency": lat,
            "decode_stats": {"bp_converged_frac": conv_all / shots_all, "osd_invocation_frac": osd_all / shots_all,
                             "logical_failures_last_step": int(counters[1].item()) if have_logicals else None,
                             "shots_last_step": int(counters[0].item())},
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--p", type=float, default=None, help="bit-flip probability (default: the config's)")
    ap.add_argument("--ms-scaling-factor", type=float, default=None,
                    help="min-sum scaling (0 = 1-2^-it, the README's; 0.625 = the reference harness's default, css_decode_sim.py:71)")
    ap.add_argument("--osd-method", default=None, choices=["osd0", "osd_e", "osd_cs"])
    ap.add_argument("--osd-order", type=int, default=None)
    ap.add_argument("--max-iter", type=int, default=None)
    ap.add_argument("--precision", type=int, default=64, choices=[64, 32])
    ap.add_argument("--shots-per-gpu", type=int, default=0)
    ap.add_argument("--e2e-shots-per-gpu", type=int, default=0)
    ap.add_argument("--cpu-shots-per-core", type=int, default=0)
    ap.add_argument("--bp-kernel", type=int, default=None)
    ap.add_argument("--bp-threads", type=int, default=0)
    ap.add_argument("--osd-variant", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    c = dict(CONFIGS[args.config])
    if args.p is not None:
        c["p"] = args.p
    if args.ms_scaling_factor is not None:
        c["ms_scaling_factor"] = args.ms_scaling_factor
    if args.osd_method is not None:
        c["osd_method"] = args.osd_method
    if args.osd_order is not None:
        c["osd_order"] = args.osd_order
    if args.max_iter is not None:
        c["max_iter"] = args.max_iter
    if c["osd_method"] == "osd0":
        c["osd_order"] = 0
    if args.impl == "reference":
        run_reference(args, args.config, c)
    else:
        run_gpu(args, args.config, c)


if __name__ == "__main__":
    main()

"""Throughput, fp32-vs-fp64 mismatch rate and logical error rates of the five BASELINE configurations.

Run on a B200 (gpurun); writes gpurun_out/<tag>_config_sweep.json and a markdown table.  The CPU column is
the oracle port on ONE core over a small sample of the same shots (test infrastructure, for context only).
"""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bp_osd_b200 import codes, BpOsdDecoder

MS = dict(max_iter=0, bp_method="ms", ms_scaling_factor=0, osd_method="osd_cs", osd_order=7)
CASES = [
    dict(cfg=1, p=0.05, shots=10_000, kw=MS, cpu=2000),
    *[dict(cfg=2, p=p, shots=1_000_000, kw=MS, cpu=1000) for p in (0.03, 0.04, 0.05, 0.06, 0.07, 0.08)],
    dict(cfg=3, p=0.05, shots=1_000_000, kw=MS, cpu=400),
    dict(cfg=4, p=0.05, shots=200_000, kw=dict(max_iter=0, bp_method="ps", ms_scaling_factor=0, osd_method="osd_e", osd_order=10), cpu=200),
    dict(cfg=5, p=0.02, shots=20_000, kw=dict(max_iter=0, bp_method="ms", ms_scaling_factor=0, osd_method="osd0", osd_order=0), cpu=20),
]


def wilson(k, n, z=1.96):
    if n == 0:
        return (0.0, 1.0)
    ph = k / n
    d = 1 + z * z / n
    c = ph + z * z / (2 * n)
    h = z * math.sqrt(ph * (1 - ph) / n + z * z / (4 * n * n))
    return ((c - h) / d, (c + h) / d)


def run_case(case, seed=0xB905D):
    code = codes.config_code(case["cfg"])
    H = code.hz
    m, n = H.shape
    shots, p = case["shots"], case["p"]
    have_l = getattr(code, "lz", None) is not None and code.lz.shape[0] > 0
    out = dict(cfg=case["cfg"], p=p, m=m, n=n, E=int(H.nnz), shots=shots, decoder=case["kw"])
    res = {}
    for prec in (64, 32):
        d = BpOsdDecoder(H, error_rate=p, precision=prec, **case["kw"])
        d.set_error_channel(px=p)
        if have_l:
            d.set_logicals(code.lz)
        err, syn = d.sample_syndromes(seed, 0, shots, sector=0)
        d.decode_batch(syn[: min(shots, 2000)], return_llr=False)          # warm-up
        torch.cuda.synchronize()
        wall = None
        dev = syn.device
        bufs = {"osdw": torch.empty((shots, n), dtype=torch.uint8, device=dev), "osd0": torch.empty((shots, n), dtype=torch.uint8, device=dev),
                "bp": torch.empty((shots, n), dtype=torch.uint8, device=dev), "converge": torch.empty(shots, dtype=torch.uint8, device=dev),
                "iter": torch.empty(shots, dtype=torch.int32, device=dev)}
        walls = []
        for _rep in range(3):   # the first full-size call also sizes the workspaces; outputs are preallocated
            torch.cuda.synchronize()
            t = time.perf_counter()
            r = d.decode_batch(syn, return_llr=False, out=bufs)
            torch.cuda.synchronize()
            walls.append(time.perf_counter() - t)
        wall = min(walls)
        st, info = d.stats(), d.info()
        fails = int(d.logical_check(err, r.osdw_decoding).sum()) if have_l else None
        res[prec] = dict(r=r.osdw_decoding, fails=fails)
        lo, hi = wilson(fails, shots) if have_l else (None, None)
        out[f"fp{prec}"] = dict(
            shots_per_s=shots / wall, wall_ms=wall * 1e3, wall_ms_reps=[w * 1e3 for w in walls], bp_ms=st["ms_bp"], osd_ms=st["ms_osd"],
            bp_shot_iterations_per_s=st["bp_iterations"] / (st["ms_bp"] * 1e-3) if st["ms_bp"] else None,
            mean_iterations=st["bp_iterations"] / shots, bp_converged_frac=st["bp_converged"] / shots,
            osd_invocations=st["osd_invocations"], bp_kernel=info["bp_kernel"], bp_threads=info["bp_threads"],
            bp_cluster_size=info["bp_cluster_size"], osd_variant=info["osd_variant"],
            logical_failures=fails, ler=(fails / shots if have_l else None), ler_ci95=[lo, hi])
    mism = int((res[64]["r"] != res[32]["r"]).any(1).sum())
    out["fp32_vs_fp64_decoding_mismatch_rate"] = mism / shots
    if have_l:
        lo, hi = out["fp64"]["ler_ci95"]
        out["fp32_ler_within_fp64_ci95"] = bool(lo <= out["fp32"]["ler"] <= hi)
    # CPU port, one core, small sample of the same syndromes
    if case.get("cpu"):
        from oracle.oracle import OracleDecoder
        o = OracleDecoder(H, error_rate=p, **case["kw"])
        s = syn[: case["cpu"]].cpu().numpy()
        t = time.perf_counter()
        ref = o.decode_batch(s, want_llr=False)
        dt = time.perf_counter() - t
        out["cpu_port_one_core"] = dict(shots=case["cpu"], shots_per_s=case["cpu"] / dt)
        same = (res[64]["r"][: case["cpu"]].cpu().numpy() == ref["osdw"]).all(1)
        out["fp64_equals_oracle_on_cpu_sample"] = bool(same.all())
        out["fp64_oracle_equal_fraction"] = float(same.mean())
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--cfgs", default="1,2,3,4,5")
    a = ap.parse_args()
    want = {int(x) for x in a.cfgs.split(",")}
    rows = []
    for case in CASES:
        if case["cfg"] in want:
            r = run_case(case)
            rows.append(r)
            print(json.dumps(r), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/{a.tag}_config_sweep.json", "w") as f:
        json.dump(rows, f, indent=1)
    with open(f"gpurun_out/{a.tag}_config_sweep.md", "w") as f:
        f.write("| cfg | p | shots | fp64 shots/s | fp64 BP it/s | mean it | BP conv | fp32 shots/s | fp32 mismatch | LER fp64 (95% CI) | LER fp32 | in CI | CPU port 1 core shots/s | fp64 == oracle |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            a64, a32 = r["fp64"], r["fp32"]
            ci = a64["ler_ci95"]
            f.write(f"| {r['cfg']} | {r['p']} | {r['shots']} | {a64['shots_per_s']:.3g} | {a64['bp_shot_iterations_per_s']:.3g} | "
                    f"{a64['mean_iterations']:.1f} | {a64['bp_converged_frac']:.4f} | {a32['shots_per_s']:.3g} | "
                    f"{r['fp32_vs_fp64_decoding_mismatch_rate']:.2e} | "
                    + (f"{a64['ler']:.3e} [{ci[0]:.3e}, {ci[1]:.3e}] | {a32['ler']:.3e} | {r.get('fp32_ler_within_fp64_ci95')} | " if a64["ler"] is not None else "n/a | n/a | n/a | ")
                    + f"{r.get('cpu_port_one_core', {}).get('shots_per_s', float('nan')):.3g} | {r.get('fp64_equals_oracle_on_cpu_sample')} ({r.get('fp64_oracle_equal_fraction', float('nan')):.3f}) |\n")
    print(open(f"gpurun_out/{a.tag}_config_sweep.md").read())

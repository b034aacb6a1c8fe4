"""A/B probe for tuning builds (BPOSD_LIB=<lib>): parity of the throughput and latency paths against the oracle on the
bench code, BP kernel speed, single-shot decode() latency.  One line of output per measurement."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bp_osd_b200 import codes, BpOsdDecoder

ap = argparse.ArgumentParser()
ap.add_argument("--tag", default=os.path.basename(os.environ.get("BPOSD_LIB", "default")))
ap.add_argument("--shots", type=int, default=300000)
ap.add_argument("--no-parity", action="store_true")
ap.add_argument("--prec", type=int, nargs="*", default=[64, 32])
a = ap.parse_args()
H = codes.config_code(3, logicals=False).hz
m, n = H.shape
KW = dict(max_iter=0, bp_method="ms", ms_scaling_factor=0, osd_method="osd_cs", osd_order=7)


def cached_oracle(name, p, B, kw, seed):
    path = f"/tmp/ab_oracle_{name}.npz"
    rng = np.random.default_rng(seed)
    e = (rng.random((B, n)) < p).astype(np.uint8)
    s = np.asarray((H @ e.T) % 2, dtype=np.uint8).T.copy()
    if os.path.exists(path):
        return s, dict(np.load(path))
    from oracle.oracle import OracleDecoder
    ref = OracleDecoder(H, error_rate=p, **kw).decode_batch(s)
    np.savez(path, **ref)
    return s, ref


def mismatches(r, ref, llr_bits=True):
    g = lambda x: x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)
    mm = {k: int((g(getattr(r, k + "_decoding")) != ref[k]).any(1).sum()) for k in ("osdw", "osd0", "bp")}
    mm["conv"] = int((g(r.converge).astype(bool) != ref["converge"].astype(bool)).sum())
    mm["iter"] = int((g(r.iter) != ref["iter"]).sum())
    if llr_bits:
        mm["llr"] = int((g(r.log_prob_ratios).view(np.uint64) != ref["llr"].view(np.uint64)).any(1).sum())
    return mm


if not a.no_parity:
    s1, ref1 = cached_oracle("cfg3_p05", 0.05, 1500, KW, 11)
    kw2 = dict(KW, max_iter=30)
    s2, ref2 = cached_oracle("cfg3_p06_it30", 0.06, 96, kw2, 12)
    d = BpOsdDecoder(H, error_rate=0.05, **KW)
    r = d.decode_batch(torch.tensor(s1, device="cuda"))
    print(a.tag, "parity throughput cfg3 1500 shots:", mismatches(r, ref1), "nonconv", int((ref1["converge"] == 0).sum()), flush=True)
    r = d.decode_batch(s1[:40])  # host batch small enough for the latency path
    print(a.tag, "parity latency-path B=40:", mismatches(r, {k: v[:40] for k, v in ref1.items()}), "launches", d.stats()["launches"], flush=True)
    d2 = BpOsdDecoder(H, error_rate=0.06, **kw2)
    r = d2.decode_batch(s2[:32])
    print(a.tag, "parity latency-path OSD-heavy B=32:", mismatches(r, {k: v[:32] for k, v in ref2.items()}),
          "nonconv", int((ref2["converge"][:32] == 0).sum()), "stats", d2.stats()["osd_invocations"], flush=True)
    bad = 0
    for i in range(64):
        out = d2.decode(s2[i])
        bad += int((out != ref2["osdw"][i]).any() or (d2.osd0_decoding != ref2["osd0"][i]).any() or bool(d2.converge) != bool(ref2["converge"][i])
                   or d2.iter != ref2["iter"][i] or (d2.log_prob_ratios.view(np.uint64) != ref2["llr"][i].view(np.uint64)).any())
    print(a.tag, "parity single-shot decode() x64 (OSD-heavy): mismatching shots", bad, flush=True)

for prec in a.prec:
    d = BpOsdDecoder(H, error_rate=0.05, precision=prec, **KW)
    d.set_error_channel(px=0.05)
    _, syn = d.sample_syndromes(1, 0, a.shots, return_errors=False)
    info = d.info()
    best = None
    for rep in range(3):
        d.decode_batch(syn, return_llr=False)
        torch.cuda.synchronize()
        st = d.stats()
        if best is None or st["ms_bp"] < best["ms_bp"]:
            best = st
    print(f"{a.tag} speed fp{prec} T={info['bp_threads']} occ={info['bp_ctas_per_sm']} bp={best['ms_bp']:.1f}ms osd={best['ms_osd']:.1f}ms "
          f"it/s={best['bp_iterations']/best['ms_bp']/1e3:.1f}M shots/s(bp)={a.shots/best['ms_bp']/1e3:.3f}M mean_it={best['bp_iterations']/a.shots:.1f}", flush=True)
    s1 = syn[:400].cpu().numpy()
    for B in (1, 8, 32):
        ts = []
        its = []
        for i in range(0, 400 - B + 1, B):
            x = s1[i] if B == 1 else s1[i:i + B]
            t = time.perf_counter()
            if B == 1:
                d.decode(x)
            else:
                d.decode_batch(x)
            ts.append(time.perf_counter() - t)
            its.append(d.iter if B == 1 else 0)
        ts = np.array(ts[len(ts) // 8:]) * 1e6
        print(f"{a.tag} latency fp{prec} B={B}: p50={np.percentile(ts, 50):.1f}us p90={np.percentile(ts, 90):.1f}us p99={np.percentile(ts, 99):.1f}us "
              f"median_iter={np.median(its):.0f} n={ts.size}", flush=True)

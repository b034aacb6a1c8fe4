"""Where the single-shot latency goes (bench code, fp64): raw C-ABI call vs Python decode(), and the fixed part
(max_iter = 1, OSD off) against the per-pass part."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bp_osd_b200 import codes, BpOsdDecoder, _capi

H = codes.config_code(3, logicals=False).hz
m, n = H.shape
tag = os.path.basename(os.environ.get("BPOSD_LIB", "main"))


def pct(ts):
    ts = np.array(ts[len(ts) // 8:]) * 1e6
    return f"p50={np.percentile(ts, 50):.1f}us p90={np.percentile(ts, 90):.1f}us p99={np.percentile(ts, 99):.1f}us"


base = BpOsdDecoder(H, error_rate=0.05, max_iter=0, bp_method="ms", ms_scaling_factor=0, osd_method="osd_cs", osd_order=7)
base.set_error_channel(px=0.05)
_, syn = base.sample_syndromes(1, 0, 600, return_errors=False)
s = syn.cpu().numpy()
for prec in (64, 32):
    for name, kw in (("full", dict(max_iter=0, osd_method="osd_cs", osd_order=7)),):
        d = BpOsdDecoder(H, error_rate=0.05, bp_method="ms", ms_scaling_factor=0, precision=prec, **kw)
        for i in range(50):
            d.decode(s[i])
        ts = []
        for i in range(600):
            t = time.perf_counter(); d.decode(s[i]); ts.append(time.perf_counter() - t)
        print(tag, f"fp{prec}", name, "python decode():", pct(ts), flush=True)
        one = d._one
        fn, ptr, h = one["fn"], one["args"], d._h
        ts, its = [], []
        for i in range(600):
            one["synd"][0] = s[i]
            t = time.perf_counter(); fn(h, ptr[0], 1, ptr[1], ptr[2], ptr[3], ptr[4], ptr[5], ptr[6]); ts.append(time.perf_counter() - t)
            its.append(int(one["iter"][0]))
        print(tag, f"fp{prec}", name, "raw bposd_decode_host B=1 (all outputs):", pct(ts), flush=True)
        x, y = np.array(its[80:], float), np.array(ts[80:]) * 1e6
        keep = (x < 1000) & (one["converge"][0] >= 0)
        b, a0 = np.polyfit(x[keep], y[keep], 1)
        print(tag, f"fp{prec}", name, f"fit: latency = {a0:.1f} us + {b:.3f} us x passes   (median passes {np.median(x):.0f}, n={keep.sum()})", flush=True)
        ts = []
        for i in range(600):
            one["synd"][0] = s[i]
            t = time.perf_counter(); fn(h, ptr[0], 1, ptr[1], None, None, None, ptr[5], ptr[6]); ts.append(time.perf_counter() - t)
        print(tag, f"fp{prec}", name, "raw bposd_decode_host B=1 (osdw, converge, iter only):", pct(ts), flush=True)

"""Ad-hoc GPU check: parity of every BP kernel variant against the oracle on small batches, plus timing."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bp_osd_b200 import codes, BpOsdDecoder
from oracle.oracle import OracleDecoder

def synd(H, p, B, rng):
    e = (rng.random((B, H.shape[1])) < p).astype(np.uint8)
    return e, (e @ H.T.toarray() % 2).astype(np.uint8)

def compare(name, H, p, B, kw, kernels=(None, 0, 1, 2), prec=64):
    rng = np.random.default_rng(12345)
    e, s = synd(H, p, B, rng)
    o = OracleDecoder(H, error_rate=p, **kw)
    t = time.time(); ref = o.decode_batch(s); tor = time.time() - t
    for k in kernels:
        d = BpOsdDecoder(H, error_rate=p, precision=prec, **kw)
        try:
            d.set_tuning(bp_kernel=k)
        except Exception as ex:
            print(name, "kernel", k, "unavailable:", ex); continue
        info = d.info()
        if k is not None and info["bp_kernel"] != k:
            print(name, "kernel", k, "-> fell to", info["bp_kernel"], "skip"); continue
        st = torch.tensor(s, device="cuda")
        torch.cuda.synchronize(); t = time.time()
        r = d.decode_batch(st)
        torch.cuda.synchronize(); tg = time.time() - t
        stt = d.stats()
        mm = {
            "osdw": int((r.osdw_decoding.cpu().numpy() != ref["osdw"]).any(1).sum()),
            "osd0": int((r.osd0_decoding.cpu().numpy() != ref["osd0"]).any(1).sum()),
            "bp": int((r.bp_decoding.cpu().numpy() != ref["bp"]).any(1).sum()),
            "conv": int((r.converge.cpu().numpy() != ref["converge"].astype(bool)).sum()),
            "iter": int((r.iter.cpu().numpy() != ref["iter"]).sum()),
        }
        llr = r.log_prob_ratios.cpu().numpy().astype(np.float64)
        if prec == 64:
            mm["llr_bits"] = int((llr != ref["llr"]).any(1).sum())
        with np.errstate(invalid="ignore", divide="ignore"):
            rel = np.nanmax(np.abs(llr - ref["llr"]) / np.maximum(np.abs(ref["llr"]), 1e-300))
        print(f"{name} prec{prec} kernel={info['bp_kernel']} T={info['bp_threads']} occ={info['bp_ctas_per_sm']} smem={info['bp_smem_bytes']} "
              f"B={B} mismatches={mm} llr_rel={rel:.2e} conv={stt['bp_converged']} osd={stt['osd_invocations']} "
              f"its={stt['bp_iterations']} ms_bp={stt['ms_bp']:.2f} ms_osd={stt['ms_osd']:.2f} wall={tg*1e3:.1f}ms oracle={tor*1e3:.0f}ms", flush=True)

if __name__ == "__main__":
    which = sys.argv[1:] or ["1", "2", "3", "4"]
    ms = dict(max_iter=0, bp_method="ms", ms_scaling_factor=0, osd_method="osd_cs", osd_order=7)
    if "1" in which:
        compare("cfg1", codes.config_code(1).hz, 0.05, 2000, ms)
        compare("cfg1-e", codes.config_code(1).hz, 0.1, 2000, dict(max_iter=7, bp_method="ms", ms_scaling_factor=0.625, osd_method="osd_e", osd_order=9))
    if "2" in which:
        compare("cfg2", codes.config_code(2).hz, 0.05, 2000, ms)
    if "3" in which:
        compare("cfg3", codes.config_code(3).hz, 0.05, 1000, ms)
        compare("cfg3-it30", codes.config_code(3).hz, 0.06, 300, dict(ms, max_iter=30))
    if "4" in which:
        compare("cfg4", codes.config_code(4).hz, 0.05, 300, dict(max_iter=0, bp_method="ps", ms_scaling_factor=0, osd_method="osd_e", osd_order=10), kernels=(None, 0, 1))
    if "f32" in which:
        compare("cfg3", codes.config_code(3).hz, 0.05, 1000, ms, prec=32)
    if "big" in which:
        d = BpOsdDecoder(codes.config_code(3).hz, error_rate=0.05, **ms)
        for B in (20000, 200000):
            rng = np.random.default_rng(1)
            e, s = synd(codes.config_code(3).hz, 0.05, B, rng)
            st = torch.tensor(s, device="cuda")
            for _ in range(2):
                torch.cuda.synchronize(); t = time.time()
                r = d.decode_batch(st, return_llr=False)
                torch.cuda.synchronize(); tg = time.time() - t
                stt = d.stats()
                print("cfg3 B", B, "wall", tg, "shots/s", B / tg, stt, flush=True)

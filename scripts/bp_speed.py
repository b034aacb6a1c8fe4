"""Quick BP-kernel speed probe on the bench workload (device sampler, events from the library)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bp_osd_b200 import codes, BpOsdDecoder

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", type=int, default=3)
ap.add_argument("--p", type=float, default=0.05)
ap.add_argument("--prec", type=int, default=64)
ap.add_argument("--shots", type=int, default=300000)
ap.add_argument("--kernel", type=int, default=None)
ap.add_argument("--threads", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--osd", default="osd_cs")
ap.add_argument("--order", type=int, default=7)
ap.add_argument("--max-iter", type=int, default=0)
ap.add_argument("--method", default="ms")
ap.add_argument("--osd-variant", type=int, default=None)
ap.add_argument("--llr", action="store_true")
ap.add_argument("--alpha", type=float, default=0.0)
a = ap.parse_args()
code = codes.config_code(a.cfg, logicals=False) if a.cfg == 4 else codes.config_code(a.cfg)
H = code.hz
d = BpOsdDecoder(H, error_rate=a.p, max_iter=a.max_iter, bp_method=a.method, ms_scaling_factor=a.alpha, osd_method=a.osd,
                 osd_order=a.order, precision=a.prec)
if a.kernel is not None or a.threads:
    d.set_tuning(bp_kernel=a.kernel, bp_threads=a.threads)
if a.osd_variant is not None:
    d.set_osd_variant(a.osd_variant)
d.set_error_channel(px=a.p)
_, syn = d.sample_syndromes(1, 0, a.shots, return_errors=False)
info = d.info()
for rep in range(a.reps):
    torch.cuda.synchronize(); t = time.time()
    d.decode_batch(syn, return_llr=a.llr)
    torch.cuda.synchronize(); wall = time.time() - t
    st = d.stats()
    print(f"cfg{a.cfg} fp{a.prec} k={info['bp_kernel']} T={info['bp_threads']} occ={info['bp_ctas_per_sm']} smem={info['bp_smem_bytes']} "
          f"shots={a.shots} wall={wall*1e3:.1f}ms bp={st['ms_bp']:.1f}ms osd={st['ms_osd']:.1f}ms "
          f"it/s={st['bp_iterations']/st['ms_bp']/1e3:.1f}M shots/s(bp)={a.shots/st['ms_bp']/1e3:.3f}M mean_it={st['bp_iterations']/a.shots:.1f} "
          f"conv={st['bp_converged']/a.shots:.4f} osd={st['osd_invocations']} chunks={st['chunks']} layout_excess={info['bp_layout_excess']} osdv={info['osd_variant']}", flush=True)

"""Cluster-kernel probe: iterations/s of BP kernel 3 by cluster size on a mid-size code, against kernel 2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bp_osd_b200 import codes, BpOsdDecoder
from bp_osd_b200.hgp import hgp

code = hgp(codes.regular_ldpc(36, 48, 3, 4, seed=11), compute_logicals=False)
H = code.hz
p, shots = 0.06, 20000
for prec in (64, 32):
    for kernel, cl in ((2, 0), (3, 2), (3, 4), (3, 8), (3, 16)):
        d = BpOsdDecoder(H, error_rate=p, max_iter=30, bp_method="ms", ms_scaling_factor=0, osd_method="off", precision=prec)
        try:
            d.set_tuning(bp_kernel=kernel)
            if kernel == 3:
                d.set_cluster_size(cl)
        except Exception as ex:
            print(f"fp{prec} kernel {kernel} cl {cl}: unavailable ({ex})"); continue
        d.set_error_channel(px=p)
        _, syn = d.sample_syndromes(1, 0, shots, return_errors=False)
        info = d.info()
        for rep in range(2):
            d.decode_batch(syn, return_llr=False); torch.cuda.synchronize()
        st = d.stats()
        ncl = max(1, info["bp_ctas_per_sm"] * info["sm_count"]) if kernel == 2 else None
        print(f"fp{prec} k={info['bp_kernel']} CL={info['bp_cluster_size']} T={info['bp_threads']} smem={info['bp_smem_bytes']} "
              f"remote_permille={info['bp_layout_excess']} bp={st['ms_bp']:.1f}ms its={st['bp_iterations']} "
              f"it/s={st['bp_iterations']/st['ms_bp']/1e3:.2f}M", flush=True)

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bp_osd_b200 import codes, BpOsdDecoder
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from _util import random_syndromes
H = codes.config_code(2).hz
kw = dict(max_iter=0, bp_method="ms", ms_scaling_factor=0, osd_method="osd_cs", osd_order=7)
_, syn = random_syndromes(H, 0.06, 1500, seed=31)
s = torch.tensor(syn, device="cuda")
ref = BpOsdDecoder(H, error_rate=0.06, precision=32, **kw); ref.set_tuning(bp_kernel=2)
r2 = ref.decode_batch(s)
for mi in (1, 2, 3, 5, 0):
    kw2 = dict(kw, max_iter=mi)
    a = BpOsdDecoder(H, error_rate=0.06, precision=32, **kw2); a.set_tuning(bp_kernel=2)
    b = BpOsdDecoder(H, error_rate=0.06, precision=32, **kw2); b.set_tuning(bp_kernel=3); b.set_cluster_size(4)
    ra, rb = a.decode_batch(s), b.decode_batch(s)
    la, lb = ra.log_prob_ratios.cpu().numpy(), rb.log_prob_ratios.cpu().numpy()
    neq = (la != lb) & ~(np.isnan(la) & np.isnan(lb))
    shots = np.flatnonzero(neq.any(1))
    print(f"max_iter={mi}: differing entries {neq.sum()} in {shots.size} shots; nan a {np.isnan(la).sum()} b {np.isnan(lb).sum()}; iter equal {(ra.iter == rb.iter).all().item()}")
    if shots.size:
        sh = shots[0]; js = np.flatnonzero(neq[sh])[:6]
        print("  shot", sh, "iter a/b", int(ra.iter[sh]), int(rb.iter[sh]), "bits", js, "a", la[sh, js], "b", lb[sh, js], "rel", np.abs(la[sh,js]-lb[sh,js])/np.abs(la[sh,js]))

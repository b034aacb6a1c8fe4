#!/bin/bash
# round 2, GPU call U: device math == host math; the cluster kernel's rows-of-7 class (cluster of 8 for config 5 on
# uniform-prior launches); cfg 5 speed; fp64 / int32 / shared-memory peaks
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "device_math or cluster or config5_full or hgp40k or golden or product_sum" --durations=5 > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2u_pytest.log
tail -8 gpurun_out/r2u_pytest.log
{
timeout 300 python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 32768 --reps 2 --osd osd0
timeout 300 python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 32768 --reps 2 --osd osd0 --prec 32
python - <<'PY'
import torch
from bp_osd_b200 import codes, BpOsdDecoder
H = codes.config_code(5).hz
for cl in (16, 8):
    d = BpOsdDecoder(H, error_rate=0.02, max_iter=0, bp_method="ms", ms_scaling_factor=0, osd_method="osd0")
    try:
        d.set_cluster_size(cl)
    except Exception as ex:
        print("cluster size", cl, "unavailable:", ex); continue
    d.set_error_channel(px=0.02)
    _, syn = d.sample_syndromes(1, 0, 16384, return_errors=False)
    info = d.info()
    for rep in range(2):
        d.decode_batch(syn, return_llr=False); torch.cuda.synchronize()
        st = d.stats()
        print(f"cfg5 forced CL={info['bp_cluster_size']} T={info['bp_threads']} smem={info['bp_smem_bytes']} remote_permille={info['bp_layout_excess']} "
              f"bp={st['ms_bp']:.1f}ms osd={st['ms_osd']:.1f}ms it/s={st['bp_iterations']/st['ms_bp']/1e3:.3f}M", flush=True)
d = BpOsdDecoder(codes.config_code(1).hz, error_rate=0.05)
print("fp64 peak %.3f T DFMA/s, int32 peak %.3f T LOP3/s, smem peak %.2f TB/s" % (d.fp64_peak() / 1e12, d.int32_peak() / 1e12, d.smem_peak() / 1e12))
PY
} > gpurun_out/r2u_speed.log 2>&1
cat gpurun_out/r2u_speed.log

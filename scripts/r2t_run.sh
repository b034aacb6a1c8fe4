#!/bin/bash
# round 2, GPU call T: device side of bposd_math.h against the host side (in-range division sequence, straight-line tanh),
# product-sum parity, and the product-sum kernel at register caps 56 / 64 (default) / 80 / 96
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "device_math" > gpurun_out/r2t_pytest.log 2>&1; tail -5 gpurun_out/r2t_pytest.log
{
for lib in "" ab/lib_ps56.so ab/lib_ps80.so ab/lib_ps96.so; do
  echo "== BPOSD_LIB=$lib"
  export BPOSD_LIB=$lib; [ -z "$lib" ] && unset BPOSD_LIB
  timeout 300 python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 200000 --reps 2 | tail -1
done
unset BPOSD_LIB
} > gpurun_out/r2t_ps_ab.log 2>&1
cat gpurun_out/r2t_ps_ab.log
python - <<'PY' 2>&1 | tee gpurun_out/r2t_peaks.log
from bp_osd_b200 import codes, BpOsdDecoder
d = BpOsdDecoder(codes.config_code(1).hz, error_rate=0.05)
print("fp64 peak %.3f T DFMA/s, int32 peak %.3f T LOP3/s, smem peak %.2f TB/s" % (d.fp64_peak() / 1e12, d.int32_peak() / 1e12, d.smem_peak() / 1e12))
PY

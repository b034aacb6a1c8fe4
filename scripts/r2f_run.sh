#!/bin/bash
# round 2, GPU call F: large-sample parity + new goldens, fp64 geometry A/B, ncu of the cluster kernel (cfg 5) and the product-sum kernel (cfg 4)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity_large.py "tests/test_gpu_parity.py::test_golden_fixtures" -x -q --durations=8 > gpurun_out/r2f_pytest_large.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2f_pytest_large.log
tail -15 gpurun_out/r2f_pytest_large.log
: > gpurun_out/r2f_ab_probe.log
timeout 400 python scripts/ab_probe.py --tag default --prec 64 >> gpurun_out/r2f_ab_probe.log 2>&1
for v in r64 v4t512; do
  BPOSD_LIB=ab/lib_$v.so timeout 400 python scripts/ab_probe.py --tag $v --prec 64 >> gpurun_out/r2f_ab_probe.log 2>&1 || echo "$v FAILED rc=$?" >> gpurun_out/r2f_ab_probe.log
done
cat gpurun_out/r2f_ab_probe.log
bash scripts/r2_ncu.sh r2f_cluster bp_cluster python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 64 --reps 1 --max-iter 150 --osd osd0
bash scripts/r2_ncu.sh r2f_ps bp_generic python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 20000 --reps 1
ls -la gpurun_out

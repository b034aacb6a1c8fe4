#!/bin/bash
# round 2, GPU call O: product-sum register caps 64 / 48 / 40, f4 test, ncu of the cluster OSD-0 kernel (second launch: the one that takes the chunk)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "received_vector or product_sum" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2o_pytest.log; tail -4 gpurun_out/r2o_pytest.log
{
for lib in "" ab/lib_ps48.so ab/lib_ps40.so; do
  echo "== BPOSD_LIB=$lib"
  export BPOSD_LIB=$lib; [ -z "$lib" ] && unset BPOSD_LIB
  timeout 300 python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 200000 --reps 2 | tail -1
done
} > gpurun_out/r2o_ps_ab.log 2>&1
cat gpurun_out/r2o_ps_ab.log
unset BPOSD_LIB
CMD="python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 64 --reps 1 --max-iter 150 --osd osd0"
ncu --set full --clock-control none --import-source on -k regex:osd0_cluster -s 1 -c 1 -o gpurun_out/r2o_osdc -f $CMD > gpurun_out/r2o_osdc_ncu.log 2>&1
tail -2 gpurun_out/r2o_osdc_ncu.log
ncu -i gpurun_out/r2o_osdc.ncu-rep --page raw --csv > gpurun_out/r2o_osdc_raw.csv 2>/dev/null
ncu -i gpurun_out/r2o_osdc.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/r2o_osdc_source.csv.gz
rm -f gpurun_out/r2o_osdc.ncu-rep

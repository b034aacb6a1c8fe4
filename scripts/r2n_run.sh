#!/bin/bash
# round 2, GPU call N: product-sum in-place kernel, register caps 128 / 80 / 64 (resident warps vs spills) + ncu of the default
mkdir -p gpurun_out
{
for lib in "" ab/lib_ps80.so ab/lib_ps64.so; do
  echo "== BPOSD_LIB=$lib"
  export BPOSD_LIB=$lib; [ -z "$lib" ] && unset BPOSD_LIB
  timeout 300 python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 200000 --reps 2 | tail -1
done
} > gpurun_out/r2n_ps_ab.log 2>&1
cat gpurun_out/r2n_ps_ab.log
unset BPOSD_LIB
bash scripts/r2_ncu.sh r2n_ps bp_fast python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 20000 --reps 1

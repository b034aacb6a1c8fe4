#!/bin/bash
# round 2, GPU call S: product-sum A/B of the two math implementations (same kernels, 64-register cap), cfg 5 with the larger workspace
mkdir -p gpurun_out
{
for lib in "" ab/lib_oldmath.so; do
  echo "== BPOSD_LIB=$lib"
  export BPOSD_LIB=$lib; [ -z "$lib" ] && unset BPOSD_LIB
  timeout 300 python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 200000 --reps 2 | tail -1
  timeout 300 python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 200000 --reps 2 --kernel 1 | tail -1
done
} > gpurun_out/r2s_ps_math_ab.log 2>&1
cat gpurun_out/r2s_ps_math_ab.log
unset BPOSD_LIB
python scripts/cfg5_sweep.py --batches 4096 32768 262144 > gpurun_out/r2s_cfg5_sweep_1gpu.jsonl 2> gpurun_out/r2s_cfg5.err; tail -n 2 gpurun_out/r2s_cfg5.err; cut -c1-250 gpurun_out/r2s_cfg5_sweep_1gpu.jsonl

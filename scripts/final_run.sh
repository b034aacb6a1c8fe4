#!/bin/bash
# Round evidence under gpurun (one GPU): GPU test suite, irregular-code A/B (cfg 2), reference arm, fp64 and fp32 bench lines,
# ncu launch list + full capture of the BP kernel.  Usage: scripts/final_run.sh <tag>
set -u
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
LOG=gpurun_out/${TAG}_cfg2_ab.log
: > $LOG
for lib in ab/lib_prev.so bp_osd_b200/libbposd_b200.so; do
  for prec in 64 32; do
    for p in 0.03 0.06; do
      echo "== $lib" >> $LOG
      BPOSD_LIB=$lib timeout 300 python scripts/bp_speed.py --cfg 2 --p $p --prec $prec --shots 1000000 --reps 2 2>&1 | tail -1 >> $LOG
    done
  done
done
cat $LOG
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python bench.py > gpurun_out/${TAG}_bench_fp64.json 2> gpurun_out/${TAG}_bench_fp64.err
python bench.py --precision 32 --no-cpu-baseline > gpurun_out/${TAG}_bench_fp32.json 2> gpurun_out/${TAG}_bench_fp32.err
scripts/profile.sh ${TAG}
cat gpurun_out/${TAG}_bench_fp64.json

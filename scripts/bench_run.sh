#!/bin/bash
# Round evidence under gpurun (one GPU): reference arm, fp64 and fp32 bench lines, then the ncu launch list and one
# full capture of the BP kernel (summarise afterwards, here: python scripts/summarize_ncu.py <tag>).  Usage: scripts/bench_run.sh <tag>
set -u
TAG=${1:-bench}
mkdir -p gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python bench.py > gpurun_out/${TAG}_bench_fp64.json 2> gpurun_out/${TAG}_bench_fp64.err
python bench.py --precision 32 --no-cpu-baseline > gpurun_out/${TAG}_bench_fp32.json 2> gpurun_out/${TAG}_bench_fp32.err
tail -c 600 gpurun_out/${TAG}_bench_fp64.err
scripts/profile.sh ${TAG}
cat gpurun_out/${TAG}_bench_fp64.json

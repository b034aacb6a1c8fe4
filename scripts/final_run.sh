#!/bin/bash
# Round evidence under gpurun (one GPU): GPU test suite, the five BASELINE configurations (scripts/config_sweep.py), reference arm, fp64 and fp32 bench lines,
# ncu launch list + full capture of the BP kernel.  Usage: scripts/final_run.sh <tag>
set -u
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python scripts/config_sweep.py --tag ${TAG} > gpurun_out/${TAG}_config_sweep.log 2>&1
tail -16 gpurun_out/${TAG}_config_sweep.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python bench.py > gpurun_out/${TAG}_bench_fp64.json 2> gpurun_out/${TAG}_bench_fp64.err
python bench.py --precision 32 --no-cpu-baseline > gpurun_out/${TAG}_bench_fp32.json 2> gpurun_out/${TAG}_bench_fp32.err
scripts/profile.sh ${TAG}
cat gpurun_out/${TAG}_bench_fp64.json

#!/bin/bash
# round 2: parity suite + bench lines (headline fp64, OSD-heavy harness-default scaling, fp32, CPU arm)
TAG=${1:-r2e}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -5 gpurun_out/${TAG}_pytest_gpu.log
python bench.py > gpurun_out/${TAG}_bench_fp64.json 2> gpurun_out/${TAG}_bench_fp64.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench_fp64.json; tail -5 gpurun_out/${TAG}_bench_fp64.err
python bench.py --ms-scaling-factor 0.625 --shots-per-gpu 200000 --steps 3 --warmup 3 --cpu-shots-per-core 150 > gpurun_out/${TAG}_bench_fp64_osdheavy.json 2> gpurun_out/${TAG}_bench_osdheavy.err; echo "bench osd-heavy rc=$?"; tail -c 2500 gpurun_out/${TAG}_bench_fp64_osdheavy.json; tail -5 gpurun_out/${TAG}_bench_osdheavy.err
python bench.py --precision 32 --no-cpu-baseline > gpurun_out/${TAG}_bench_fp32.json 2> gpurun_out/${TAG}_bench_fp32.err; echo "bench fp32 rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_fp32.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>&1; echo "ref rc=$?"; tail -c 800 gpurun_out/${TAG}_bench_reference.json

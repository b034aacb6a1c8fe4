#!/bin/bash
# round 2, GPU call V: evidence with the final library -- full GPU parity suite, bench lines (headline fp64, harness-default
# scaling, fp32, product-sum cfg 4, cfg 5, CPU arm), ncu launch list + full capture of the BP kernel, ncu of the product-sum
# kernel, cfg 5 batch sweep on one GPU
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_reference.json
python bench.py > gpurun_out/${TAG}_bench_fp64.json 2> gpurun_out/${TAG}_bench_fp64.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/${TAG}_bench_fp64.json; tail -3 gpurun_out/${TAG}_bench_fp64.err
python bench.py --ms-scaling-factor 0.625 --shots-per-gpu 200000 --steps 3 --warmup 3 --cpu-shots-per-core 150 > gpurun_out/${TAG}_bench_fp64_osdheavy.json 2> gpurun_out/${TAG}_bench_osdheavy.err; echo "osd-heavy rc=$?"; tail -c 1200 gpurun_out/${TAG}_bench_fp64_osdheavy.json
python bench.py --precision 32 --no-cpu-baseline > gpurun_out/${TAG}_bench_fp32.json 2> gpurun_out/${TAG}_bench_fp32.err; echo "fp32 rc=$?"; tail -c 1200 gpurun_out/${TAG}_bench_fp32.json
python bench.py --config 4 --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_cfg4.json 2> gpurun_out/${TAG}_bench_cfg4.err; echo "cfg4 rc=$?"; tail -c 2500 gpurun_out/${TAG}_bench_cfg4.json; tail -3 gpurun_out/${TAG}_bench_cfg4.err
python bench.py --config 5 --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5.err; echo "cfg5 rc=$?"; tail -c 1500 gpurun_out/${TAG}_bench_cfg5.json; tail -3 gpurun_out/${TAG}_bench_cfg5.err
python scripts/cfg5_sweep.py --batches 1 8 64 512 4096 32768 262144 > gpurun_out/${TAG}_cfg5_sweep_1gpu.jsonl 2> gpurun_out/${TAG}_cfg5.err; tail -n 2 gpurun_out/${TAG}_cfg5.err; cut -c1-250 gpurun_out/${TAG}_cfg5_sweep_1gpu.jsonl
python scripts/cfg5_sweep.py --batches 262144 1048576 --chunk 262144 > gpurun_out/${TAG}_cfg5_sweep_1gpu_bigchunk.jsonl 2>> gpurun_out/${TAG}_cfg5.err; cut -c1-250 gpurun_out/${TAG}_cfg5_sweep_1gpu_bigchunk.jsonl
scripts/profile.sh ${TAG}
bash scripts/r2_ncu.sh ${TAG}_ps bp_fast python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 20000 --reps 1

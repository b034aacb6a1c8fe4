#!/bin/bash
# round 2, GPU call Q (8 GPUs): cfg 3 bench at 8 ranks; cfg 5 batch sweep at 8 / 4 / 2 ranks (strong scaling over the batch)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --shots-per-gpu 200000 --no-cpu-baseline > gpurun_out/r2r_bench_2gpu_short.json 2> gpurun_out/r2r_bench_2gpu.err; echo "bench2 rc=$?"; head -c 300 gpurun_out/r2r_bench_2gpu_short.json; wc -l gpurun_out/r2r_bench_2gpu_short.json
$TR --nproc-per-node 8 --master-port 29512 scripts/cfg5_sweep.py --batches 4096 32768 262144 1048576 > gpurun_out/r2r_cfg5_sweep_8gpu.jsonl 2> gpurun_out/r2r_cfg5_8.err; echo "sweep8 rc=$?"; cat gpurun_out/r2r_cfg5_sweep_8gpu.jsonl | cut -c1-260
$TR --nproc-per-node 4 --master-port 29513 scripts/cfg5_sweep.py --batches 4096 32768 262144 > gpurun_out/r2r_cfg5_sweep_4gpu.jsonl 2> gpurun_out/r2r_cfg5_4.err; echo "sweep4 rc=$?"; cat gpurun_out/r2r_cfg5_sweep_4gpu.jsonl | cut -c1-260
$TR --nproc-per-node 2 --master-port 29514 scripts/cfg5_sweep.py --batches 4096 32768 262144 > gpurun_out/r2r_cfg5_sweep_2gpu.jsonl 2> gpurun_out/r2r_cfg5_2.err; echo "sweep2 rc=$?"; cat gpurun_out/r2r_cfg5_sweep_2gpu.jsonl | cut -c1-260
tail -n 2 gpurun_out/r2r_cfg5_8.err

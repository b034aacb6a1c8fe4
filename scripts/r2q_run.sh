#!/bin/bash
# round 2, GPU call Q (8 GPUs): cfg 3 bench at 8 ranks; cfg 5 batch sweep at 8 / 4 / 2 ranks (strong scaling over the batch)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 > gpurun_out/r2q_bench_8gpu_fp64.json 2> gpurun_out/r2q_bench_8gpu.err; echo "bench8 rc=$?"; tail -c 1200 gpurun_out/r2q_bench_8gpu_fp64.json
$TR --nproc-per-node 8 --master-port 29512 scripts/cfg5_sweep.py --batches 4096 32768 262144 1048576 > gpurun_out/r2q_cfg5_sweep_8gpu.jsonl 2> gpurun_out/r2q_cfg5_8.err; echo "sweep8 rc=$?"; cat gpurun_out/r2q_cfg5_sweep_8gpu.jsonl | cut -c1-260
$TR --nproc-per-node 4 --master-port 29513 scripts/cfg5_sweep.py --batches 4096 32768 262144 > gpurun_out/r2q_cfg5_sweep_4gpu.jsonl 2> gpurun_out/r2q_cfg5_4.err; echo "sweep4 rc=$?"; cat gpurun_out/r2q_cfg5_sweep_4gpu.jsonl | cut -c1-260
$TR --nproc-per-node 2 --master-port 29514 scripts/cfg5_sweep.py --batches 4096 32768 262144 > gpurun_out/r2q_cfg5_sweep_2gpu.jsonl 2> gpurun_out/r2q_cfg5_2.err; echo "sweep2 rc=$?"; cat gpurun_out/r2q_cfg5_sweep_2gpu.jsonl | cut -c1-260
tail -3 gpurun_out/r2q_*.err

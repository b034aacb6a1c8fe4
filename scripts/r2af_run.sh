#!/bin/bash
# round 2, GPU call AF (2 GPUs): the driver's multi-rank launch of bench.py with the final library, both arms
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2af_bench_2gpu.json 2> gpurun_out/r2af_bench_2gpu.err; echo "rc=$?"
tail -c 900 gpurun_out/r2af_bench_2gpu.json; tail -3 gpurun_out/r2af_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2af_bench_2gpu_reference.json 2> gpurun_out/r2af_bench_2gpu_reference.err; echo "ref rc=$?"
tail -c 500 gpurun_out/r2af_bench_2gpu_reference.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2af_bench_1gpu.json 2> gpurun_out/r2af_bench_1gpu.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("2gpu", "1gpu"):
    j = json.loads(open(f"gpurun_out/r2af_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "value", j["value"], "e2e", j["e2e"]["value"], "lat", j["latency"]["p50_us"], "n_gpus", j["n_gpus"])
PY

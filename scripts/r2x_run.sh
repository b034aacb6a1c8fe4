#!/bin/bash
# round 2, GPU call X: overflow guard (long runs beyond max_iter = n), rows-of-7 pad emulation, serial schedule; headline
# BP speed with the guard in place; cfg 5 sweep
mkdir -p gpurun_out
: > gpurun_out/r2x_pytest.log
for grp in "overflow" "cluster or config5 or hgp40k or standin or serial or device_math"; do
  echo "=== $grp" >> gpurun_out/r2x_pytest.log
  timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -m gpu -q -k "$grp" --durations=4 >> gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
done
grep -E "^===|^E  |passed|failed|FAILED|ERROR" gpurun_out/r2x_pytest.log | cut -c1-300 | head -40
{
timeout 300 python scripts/bp_speed.py --cfg 3 --shots 1250000 --reps 3
timeout 300 python scripts/bp_speed.py --cfg 3 --shots 1250000 --reps 2 --prec 32
timeout 300 python scripts/bp_speed.py --cfg 2 --p 0.05 --shots 1000000 --reps 2
} > gpurun_out/r2x_speed.log 2>&1
cat gpurun_out/r2x_speed.log
python scripts/cfg5_sweep.py --batches 4096 32768 262144 > gpurun_out/r2x_cfg5_sweep_1gpu.jsonl 2> gpurun_out/r2x_cfg5.err; tail -n 2 gpurun_out/r2x_cfg5.err; cut -c1-330 gpurun_out/r2x_cfg5_sweep_1gpu.jsonl

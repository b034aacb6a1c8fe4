#!/bin/bash
# round 2, GPU call AH: cluster kernel with the bit-slot offsets back in registers (flip path reads its slots from the global table)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -m gpu -q -k "cluster or config5 or hgp40k or standin or overflow" --durations=4 > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ah_pytest.log
grep -E "^E  |passed|failed|FAILED|ERROR|rc=" gpurun_out/r2ah_pytest.log | cut -c1-300 | head -20
python scripts/cfg5_sweep.py --batches 32768 262144 > gpurun_out/r2ah_cfg5_sweep_1gpu.jsonl 2> gpurun_out/r2ah_cfg5.err; tail -n 2 gpurun_out/r2ah_cfg5.err; cut -c1-330 gpurun_out/r2ah_cfg5_sweep_1gpu.jsonl

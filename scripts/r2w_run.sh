#!/bin/bash
# round 2, GPU call W: the cluster kernel's rows-of-7 class and the serial schedule against the oracle (one pytest process
# per group: a device fault in one must not hide the others)
mkdir -p gpurun_out
: > gpurun_out/r2w_pytest.log
for grp in "device_math" "serial" "cluster or config5_full or hgp40k or standin" "config5_three or config5_max"; do
  echo "=== $grp" >> gpurun_out/r2w_pytest.log
  timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -m gpu -q -k "$grp" --durations=4 >> gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log
done
grep -E "^===|^E  |passed|failed|FAILED|ERROR" gpurun_out/r2w_pytest.log | cut -c1-400 | head -60

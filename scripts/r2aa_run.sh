#!/bin/bash
# round 2, GPU call AA: cluster kernel with local parity flips by slot arithmetic + compact remote flip descriptors
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -m gpu -q -k "cluster or config5 or hgp40k or standin or aliases or overflow" --durations=4 > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aa_pytest.log
grep -E "^E  |passed|failed|FAILED|ERROR|rc=" gpurun_out/r2aa_pytest.log | cut -c1-300 | head -20
python scripts/cfg5_sweep.py --batches 4096 32768 262144 > gpurun_out/r2aa_cfg5_sweep_1gpu.jsonl 2> gpurun_out/r2aa_cfg5.err; tail -n 2 gpurun_out/r2aa_cfg5.err; cut -c1-330 gpurun_out/r2aa_cfg5_sweep_1gpu.jsonl
python - <<'PY'
from bp_osd_b200 import codes, BpOsdDecoder
d = BpOsdDecoder(codes.config_code(5).hz, error_rate=0.02, max_iter=0, bp_method="ms", ms_scaling_factor=0, osd_method="osd0")
print(d.info())
PY

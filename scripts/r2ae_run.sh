#!/bin/bash
# round 2, GPU call AE: product-sum after the log restructuring (parity + speed), then the whole GPU suite once more
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2ae_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2ae_pytest_gpu.log
tail -3 gpurun_out/r2ae_pytest_gpu.log
{
timeout 300 python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 200000 --reps 2 | tail -1
timeout 300 python scripts/bp_speed.py --cfg 2 --p 0.06 --method ps --shots 500000 --reps 2 | tail -1
} > gpurun_out/r2ae_ps_speed.log 2>&1
cat gpurun_out/r2ae_ps_speed.log
python bench.py --config 4 --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2ae_bench_cfg4.json 2> gpurun_out/r2ae_bench_cfg4.err; echo "cfg4 rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2ae_bench_cfg4.json").read().strip().splitlines()[-1]); r = j["roofline"]
print("cfg4 value", j["value"], "e2e", j["e2e"]["value"], "it/s", r["bp_shot_iterations_per_s"], "frac", r["frac"], "full-path", r["frac_full_path"])
PY

#!/bin/bash
# round 2, GPU call Y: bench lines with the overflow guard in place (like-for-like with r2v)
mkdir -p gpurun_out
python bench.py --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/r2y_bench_fp64.json 2> gpurun_out/r2y_bench_fp64.err; echo "rc=$?"
python bench.py --no-cpu-baseline --steps 4 --warmup 3 --precision 32 > gpurun_out/r2y_bench_fp32.json 2> gpurun_out/r2y_bench_fp32.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("fp64", "fp32"):
    j = json.loads(open(f"gpurun_out/r2y_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, "value", j["value"], "e2e", j["e2e"]["value"], "bp it/s", j["roofline"]["bp_shot_iterations_per_s"], "frac", j["roofline"]["frac"], "bp ms", j["roofline"]["bp_ms_per_step"])
PY

#!/bin/bash
# round 2, GPU call AD: product-sum CTA-size A/B on cfg 4 (rows per thread: 441 rows over 128 / 256 / 448 threads); headline bench with the
# logical check warmed up before the timed region
mkdir -p gpurun_out
{
for t in 0 256 448 224; do
  timeout 300 python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 200000 --reps 2 --threads $t | tail -1
done
} > gpurun_out/r2ad_ps_threads.log 2>&1
cat gpurun_out/r2ad_ps_threads.log
python bench.py --no-cpu-baseline > gpurun_out/r2ad_bench_fp64.json 2> gpurun_out/r2ad_bench_fp64.err; echo "rc=$?"
python - <<'PY'
import json
j = json.loads(open("gpurun_out/r2ad_bench_fp64.json").read().strip().splitlines()[-1]); r = j["roofline"]
print("value", j["value"], "e2e", j["e2e"]["value"], "ms/step", j["ms_per_step"], "bp", r["bp_ms_per_step"], "osd", r["osd_ms_per_step"])
PY

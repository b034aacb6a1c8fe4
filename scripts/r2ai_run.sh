#!/bin/bash
# round 2, GPU call AI: the library as committed at the end of the round -- the driver's commands (GPU suite, smoke, bench both arms)
TAG=${1:-r2ai}
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"; tail -c 300 gpurun_out/${TAG}_bench_reference.json
python bench.py > gpurun_out/${TAG}_bench_fp64.json 2> gpurun_out/${TAG}_bench_fp64.err; echo "bench rc=$?"
python - ${TAG} <<'PY'
import json
j = json.loads(open("gpurun_out/" + __import__("sys").argv[1] + "_bench_fp64.json").read().strip().splitlines()[-1]); r = j["roofline"]
print("value", j["value"], "e2e", j["e2e"]["value"], "ms/step", j["ms_per_step"], "frac", r["frac"], "lat", j["latency"]["p50_us"], "cpu", j["cpu_baseline"]["value"], "launches", j["gpu_launches"])
PY

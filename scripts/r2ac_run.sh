#!/bin/bash
# round 2, GPU call AC: the final library -- full GPU suite (no -x), headline bench, ncu launch list + full BP capture, ncu of the product-sum kernel
TAG=r2ac
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -4 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/${TAG}_bench_fp64.json 2> gpurun_out/${TAG}_bench_fp64.err; echo "bench rc=$?"; tail -c 600 gpurun_out/${TAG}_bench_fp64.json
python bench.py --config 5 --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_cfg5.json 2> gpurun_out/${TAG}_bench_cfg5.err; echo "cfg5 rc=$?"
scripts/profile.sh ${TAG}
bash scripts/r2_ncu.sh ${TAG}_ps bp_fast python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 20000 --reps 1

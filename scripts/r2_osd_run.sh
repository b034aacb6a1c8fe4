#!/bin/bash
# round 2, GPU call B: parity suite + OSD kernel probes (old shared-memory T kernel = variant 1, register kernel = 3)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2b_pytest_gpu.log
tail -15 gpurun_out/r2b_pytest_gpu.log
{
for v in 1 3; do
  python scripts/bp_speed.py --cfg 3 --max-iter 16 --shots 100000 --reps 2 --osd-variant $v
  python scripts/bp_speed.py --cfg 2 --p 0.08 --max-iter 8 --shots 300000 --reps 2 --osd-variant $v
  python scripts/bp_speed.py --cfg 4 --max-iter 8 --shots 100000 --reps 2 --osd-variant $v --osd osd_e --order 10
  python scripts/bp_speed.py --cfg 1 --p 0.1 --max-iter 3 --shots 1000000 --reps 2 --osd-variant $v
done
python scripts/bp_speed.py --cfg 3 --alpha 0.625 --shots 100000 --reps 2 --osd-variant 3
} > gpurun_out/r2b_osd_probes.log 2>&1
cat gpurun_out/r2b_osd_probes.log

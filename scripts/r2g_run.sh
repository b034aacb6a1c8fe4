#!/bin/bash
# round 2, GPU call G: full parity suite on the new cluster (halo exchange) kernel + xorsign fp32; cfg 5 / fp32 speed probes
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2g_pytest_gpu.log
tail -25 gpurun_out/r2g_pytest_gpu.log
{
python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 4000 --reps 2 --max-iter 60 --osd osd0
python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 4000 --reps 2 --max-iter 60 --osd osd0 --prec 32
python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 20000 --reps 1 --osd osd0
python scripts/bp_speed.py --cfg 3 --prec 32 --shots 300000 --reps 3
python scripts/bp_speed.py --cfg 2 --prec 32 --shots 300000 --reps 2
python scripts/bp_speed.py --cfg 2 --prec 64 --shots 300000 --reps 2
} > gpurun_out/r2g_speed.log 2>&1
cat gpurun_out/r2g_speed.log

#!/bin/bash
# round 2, GPU call M: full parity suite (product-sum on the in-place kernel, deeper OSD TMA ring) + product-sum speed
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2m_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2m_pytest_gpu.log
tail -15 gpurun_out/r2m_pytest_gpu.log
{
timeout 300 python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 200000 --reps 2
timeout 300 python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 200000 --reps 2 --kernel 1
timeout 300 python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 200000 --reps 2 --prec 32
timeout 300 python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 200000 --reps 2 --threads 512
} > gpurun_out/r2m_speed.log 2>&1
cat gpurun_out/r2m_speed.log

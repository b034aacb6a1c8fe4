#!/bin/bash
# round 2, GPU call I: cluster OSD-0 kernel (variant 4): parity on small / stand-in / full-size codes, then cfg 5 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hbm_osd0 or standin or config5_full or hgp40k" --durations=5 > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2i_pytest.log
tail -20 gpurun_out/r2i_pytest.log
{
timeout 300 python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 20000 --reps 2 --osd osd0
timeout 300 python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 20000 --reps 1 --osd osd0 --osd-variant 2
timeout 300 python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 2000 --reps 2 --max-iter 30 --osd osd0
} > gpurun_out/r2i_speed.log 2>&1
cat gpurun_out/r2i_speed.log

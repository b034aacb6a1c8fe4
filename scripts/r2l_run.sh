#!/bin/bash
# round 2, GPU call L: product-sum with the FMA branch-free tanh/log; cluster OSD-0 with the deeper TMA ring
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "product_sum or lp882 or hbm_osd0 or standin or config5_full or hgp40k or golden" --durations=5 > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2l_pytest.log
tail -8 gpurun_out/r2l_pytest.log
{
timeout 300 python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 200000 --reps 2
timeout 300 python scripts/bp_speed.py --cfg 4 --method ps --osd osd_e --order 10 --shots 200000 --reps 2 --prec 32
timeout 300 python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 32768 --reps 1 --osd osd0
timeout 300 python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 2000 --reps 1 --max-iter 30 --osd osd0
} > gpurun_out/r2l_speed.log 2>&1
cat gpurun_out/r2l_speed.log

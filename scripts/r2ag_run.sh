#!/bin/bash
# round 2, GPU call AG: ncu of the cluster BP kernel on cfg 5 with clusters of 8 (rows of 7)
mkdir -p gpurun_out
bash scripts/r2_ncu.sh r2ag_cluster bp_cluster python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 144 --reps 1 --max-iter 150 --osd osd0

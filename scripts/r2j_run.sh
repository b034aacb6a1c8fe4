#!/bin/bash
# round 2, GPU call J: ncu of the halo-exchange cluster BP kernel and of the cluster OSD-0 kernel on cfg 5
mkdir -p gpurun_out
bash scripts/r2_ncu.sh r2j_cluster bp_cluster python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 64 --reps 1 --max-iter 150 --osd osd0
bash scripts/r2_ncu.sh r2j_osdc osd0_cluster python scripts/bp_speed.py --cfg 5 --p 0.02 --shots 64 --reps 1 --max-iter 150 --osd osd0
ls -la gpurun_out

#!/bin/bash
# A/B of the OSD register kernel's round size G (default build vs ab/lib_g8.so, ab/lib_g4.so)
TAG=${1:-r2d}
mkdir -p gpurun_out
{
for lib in "" ab/lib_g8.so ab/lib_g4.so; do
  echo "== BPOSD_LIB=$lib"
  export BPOSD_LIB=$lib; [ -z "$lib" ] && unset BPOSD_LIB
  python scripts/bp_speed.py --cfg 3 --max-iter 16 --shots 100000 --reps 2 --osd-variant 3 | tail -1
  python scripts/bp_speed.py --cfg 2 --p 0.08 --max-iter 8 --shots 300000 --reps 2 --osd-variant 3 | tail -1
  python scripts/bp_speed.py --cfg 4 --max-iter 8 --shots 100000 --reps 2 --osd-variant 3 --osd osd_e --order 10 | tail -1
  python scripts/bp_speed.py --cfg 1 --p 0.1 --max-iter 3 --shots 1000000 --reps 2 --osd-variant 3 | tail -1
done
} > gpurun_out/${TAG}_osd_ab.log 2>&1
cat gpurun_out/${TAG}_osd_ab.log

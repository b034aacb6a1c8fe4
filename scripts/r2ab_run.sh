#!/bin/bash
# round 2, GPU call AB: cluster BP on cfg 5, CTA size A/B (8 bits x 640 threads vs 6 bits x 864 threads, two-pass check update)
mkdir -p gpurun_out
for v in 8 6; do
  echo "== BPOSD_CLUSTER_VPT=$v"
  BPOSD_CLUSTER_VPT=$v python scripts/cfg5_sweep.py --batches 32768 2>/dev/null | cut -c1-420
done 2>&1 | tee gpurun_out/r2ab_cluster_vpt.log

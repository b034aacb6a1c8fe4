#!/bin/bash
# ncu evidence for the bench's dominant kernel (run under gpurun; one GPU).  Usage: scripts/profile.sh <tag> [extra bench args]
set -u
TAG=${1:-r01}; shift || true
CMD="python bench.py --steps 1 --warmup 1 --shots-per-gpu 30000 --e2e-shots-per-gpu 30000 --no-cpu-baseline $*"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bp_ -s 1 -c 1 -o gpurun_out/${TAG}_bp -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out | tail -8

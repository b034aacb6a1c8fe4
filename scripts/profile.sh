#!/bin/bash
# ncu evidence for the bench's dominant kernel (run under gpurun; one GPU).  Usage: scripts/profile.sh <tag> [extra bench args]
# Leaves in gpurun_out/: <tag>_launches.csv (launch list), <tag>_bp_raw.csv (every metric of the full capture),
# <tag>_bp_source.csv.gz (per-SASS-line counters with source lines) and, if it is small enough to travel, the .ncu-rep.
set -u
TAG=${1:-r01}; shift || true
CMD="python bench.py --steps 1 --warmup 1 --shots-per-gpu 30000 --e2e-shots-per-gpu 30000 --no-cpu-baseline $*"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:bp_ -s 1 -c 1 -o gpurun_out/${TAG}_bp -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log
REP=gpurun_out/${TAG}_bp.ncu-rep
if [ -f $REP ]; then
  ncu -i $REP --page raw --csv > gpurun_out/${TAG}_bp_raw.csv 2>/dev/null
  ncu -i $REP --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${TAG}_bp_source.csv.gz
  # gpurun brings back at most 64 MiB: a report with imported sources of this library is ~100 MB
  if [ $(stat -c %s $REP) -gt 40000000 ]; then rm -f $REP; fi
fi
ls -la gpurun_out | tail -12

#!/bin/bash
# Tuning run under gpurun: every A/B library in ab/ and the release library through scripts/ab_probe.py and
# scripts/lat_probe.py, then the GPU test suite on the release library.  Usage: scripts/ab_run.sh <tag> [variants...]
set -u
TAG=${1:-ab}; shift || true
mkdir -p gpurun_out
LOG=gpurun_out/${TAG}_ab_probe.log
: > $LOG
for v in "$@"; do
  BPOSD_LIB=ab/lib_$v.so timeout 400 python scripts/ab_probe.py --tag $v >> $LOG 2>&1 || echo "$v FAILED rc=$?" >> $LOG
  BPOSD_LIB=ab/lib_$v.so timeout 300 python scripts/lat_probe.py >> $LOG 2>&1 || echo "$v FAILED rc=$?" >> $LOG
done
timeout 400 python scripts/ab_probe.py --tag main >> $LOG 2>&1 || echo "main FAILED rc=$?" >> $LOG
timeout 300 python scripts/lat_probe.py >> $LOG 2>&1 || echo "main FAILED rc=$?" >> $LOG
cat $LOG
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log

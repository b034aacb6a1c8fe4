#!/bin/bash
# Latency probes under gpurun: every library in ab/ plus the release library through scripts/lat_probe.py, then the GPU
# test suite on the release library.  Usage: scripts/lat_run.sh <tag> [variants...]
set -u
TAG=${1:-lat}; shift || true
mkdir -p gpurun_out
LOG=gpurun_out/${TAG}_lat_probe.log
: > $LOG
for v in "$@"; do
  BPOSD_LIB=ab/lib_$v.so timeout 300 python scripts/lat_probe.py >> $LOG 2>&1 || echo "$v FAILED rc=$?" >> $LOG
done
timeout 300 python scripts/lat_probe.py >> $LOG 2>&1 || echo "main FAILED rc=$?" >> $LOG
cat $LOG
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log

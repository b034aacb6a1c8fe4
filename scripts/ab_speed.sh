#!/bin/bash
# Quick A/B under gpurun: every library in ab/ named on the command line through scripts/ab_probe.py.  Usage: scripts/ab_speed.sh <tag> [variants...]
set -u
TAG=${1:-ab}; shift || true
mkdir -p gpurun_out
LOG=gpurun_out/${TAG}_ab_probe.log
: > $LOG
for v in "$@"; do
  BPOSD_LIB=ab/lib_$v.so timeout 400 python scripts/ab_probe.py --tag $v >> $LOG 2>&1 || echo "$v FAILED rc=$?" >> $LOG
done
cat $LOG

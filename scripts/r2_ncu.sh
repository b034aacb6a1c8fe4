#!/bin/bash
# ncu --set full capture (raw metrics + per-SASS source view) of one kernel of an arbitrary probe command.
# Usage: scripts/r2_ncu.sh <tag> <kernel-name regex> <command ...>      (run under gpurun, one GPU)
TAG=$1; KRE=$2; shift 2
CMD="$*"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$KRE -c 1 -o gpurun_out/${TAG} -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log; tail -2 gpurun_out/${TAG}_plain.log
REP=gpurun_out/${TAG}.ncu-rep
if [ -f $REP ]; then
  ncu -i $REP --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
  ncu -i $REP --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${TAG}_source.csv.gz
  if [ $(stat -c %s $REP) -gt 30000000 ]; then rm -f $REP; fi
fi

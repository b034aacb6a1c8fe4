#!/bin/bash
# ncu --set full capture (with SASS/source view) of the OSD kernel on the OSD-heavy cfg-3 probe
TAG=${1:-r2c}; CFG=${2:-3}; MI=${3:-16}; SHOTS=${4:-20000}
CMD="python scripts/bp_speed.py --cfg $CFG --max-iter $MI --shots $SHOTS --reps 1 --osd-variant 3"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:osd_reg -c 1 -o gpurun_out/${TAG}_osd -f $CMD > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log; cat gpurun_out/${TAG}_plain.log
REP=gpurun_out/${TAG}_osd.ncu-rep
if [ -f $REP ]; then
  ncu -i $REP --page raw --csv > gpurun_out/${TAG}_osd_raw.csv 2>/dev/null
  ncu -i $REP --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${TAG}_osd_source.csv.gz
  ls -la $REP; if [ $(stat -c %s $REP) -gt 40000000 ]; then rm -f $REP; fi
fi
ls -la gpurun_out | tail -8

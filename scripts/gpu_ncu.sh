#!/bin/bash
# ncu --set full capture of selected kernel instances of eager training steps; exports the raw metrics as CSV next to
# the report and drops reports that are too large to travel.  Usage: scripts/gpu_ncu.sh <tag> <regex> <skip> <count>
TAG=$1; RE=$2; SKIP=$3; CNT=$4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$RE" -s $SKIP -c $CNT -f -o gpurun_out/${TAG} python scripts/profile_step.py --steps 3 > gpurun_out/${TAG}_ncu.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}.raw.csv 2>/dev/null
SZ=$(stat -c %s gpurun_out/${TAG}.ncu-rep 2>/dev/null || echo 0)
if [ "$SZ" -gt 12000000 ]; then rm -f gpurun_out/${TAG}.ncu-rep; fi
tail -1 gpurun_out/${TAG}_ncu.log; ls -la gpurun_out/ | grep ${TAG}

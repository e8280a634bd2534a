#!/bin/bash
# One GPU round trip: parity tests, headline bench, ncu launch list of one eager step.  Usage: scripts/gpu_cycle.sh <tag> [pytest args]
TAG=${1:-run}; shift
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -15) > gpurun_out/${TAG}_tests.log
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.log 2>&1
LPS=$(timeout 300 python scripts/profile_step.py --steps 1 2>/dev/null | awk '/launches_per_step/{print $2}')
echo "launches_per_step=$LPS" > gpurun_out/${TAG}_prof.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2*LPS)) -c $LPS --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/profile_step.py --steps 3 >> gpurun_out/${TAG}_prof.log 2>&1
tail -4 gpurun_out/${TAG}_tests.log; tail -1 gpurun_out/${TAG}_bench.log | cut -c1-400; tail -3 gpurun_out/${TAG}_prof.log

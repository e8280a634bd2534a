LPS=$(timeout 300 python scripts/profile_step.py --steps 1 2>/dev/null | awk '/launches_per_step/{print $2}')
echo "launches_per_step=$LPS" > gpurun_out/s28_prof.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2*LPS)) -c $LPS --csv --log-file gpurun_out/s28_launches.csv python scripts/profile_step.py --steps 3 >> gpurun_out/s28_prof.log 2>&1
tail -3 gpurun_out/s28_prof.log

LPS=$(timeout 300 python scripts/profile_step.py --config people --steps 1 2>/dev/null | awk '/launches_per_step/{print $2}')
echo "launches_per_step=$LPS"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2*LPS)) -c $LPS --csv --log-file gpurun_out/r2_people_launches.csv python scripts/profile_step.py --config people --steps 3 > gpurun_out/r2_people_prof.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_people_launches.csv | head -30
for spec in "dws1:dwconv_s1_kernel:6:2" "dwwg:dwconv_wgrad_kernel:6:1"; do
  IFS=: read NAME RE SKIP CNT <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$RE" -s $SKIP -c $CNT -f -o gpurun_out/r2_people_${NAME} python scripts/profile_step.py --config people --steps 1 > gpurun_out/r2_people_${NAME}.log 2>&1
  ncu -i gpurun_out/r2_people_${NAME}.ncu-rep --page raw --csv > gpurun_out/r2_people_${NAME}.raw.csv 2>/dev/null
  python scripts/ncu_raw.py gpurun_out/r2_people_${NAME}.raw.csv > gpurun_out/r2_ncu_full_people_${NAME}.metrics.txt 2>/dev/null
  echo "== $NAME"; grep -E "gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |sm__throughput.avg.pct|smsp__inst_executed.sum |warps_active.avg.pct|l1tex__throughput.avg.pct|lts__throughput.avg.pct|launch__grid_size|registers_per_thread" gpurun_out/r2_ncu_full_people_${NAME}.metrics.txt | head -20
done

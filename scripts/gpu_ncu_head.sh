for spec in "headw2:head_wgrad_strip_kernel" "headd2:head_dgrad1_strip_kernel"; do
  IFS=: read NAME RE <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$RE" -s 1 -c 1 -f -o gpurun_out/r2_tail_${NAME} python scripts/profile_step.py --steps 2 > gpurun_out/r2_tail_${NAME}.log 2>&1
  ncu -i gpurun_out/r2_tail_${NAME}.ncu-rep --page raw --csv > gpurun_out/r2_tail_${NAME}.raw.csv 2>/dev/null
  python scripts/ncu_raw.py gpurun_out/r2_tail_${NAME}.raw.csv > gpurun_out/r2_ncu_full_tail_${NAME}.metrics.txt 2>/dev/null
  echo "== $NAME"; grep -E "gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |sm__throughput.avg.pct|smsp__inst_executed.sum |warps_active.avg.pct|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate.pct|registers_per_thread|launch__grid_size|l1tex__throughput.avg.pct|lts__throughput.avg.pct|smsp__cycles_active.avg |stall" gpurun_out/r2_ncu_full_tail_${NAME}.metrics.txt | head -40
done

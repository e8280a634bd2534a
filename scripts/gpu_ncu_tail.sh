mkdir -p gpurun_out
for spec in "headw:head_wgrad_kernel:0:1" "headd:head_dgrad1_kernel:0:1" "mpb:maxpool_bwd_k3s2_kernel:0:1" "mpf:maxpool_fwd_kernel:0:1" "aug:augment_apply_kernel:0:1" "wprep:weight_prep_batched_kernel:0:1"; do
  IFS=: read NAME RE SKIP CNT <<< "$spec"
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$RE" -s 1 -c 1 -f -o gpurun_out/r2_tail_${NAME} python scripts/profile_step.py --steps 2 > gpurun_out/r2_tail_${NAME}.log 2>&1
  ncu -i gpurun_out/r2_tail_${NAME}.ncu-rep --page raw --csv > gpurun_out/r2_tail_${NAME}.raw.csv 2>/dev/null
  python scripts/ncu_raw.py gpurun_out/r2_tail_${NAME}.raw.csv > gpurun_out/r2_ncu_full_tail_${NAME}.metrics.txt 2>/dev/null
  echo "== $NAME"; grep -E "gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |sm__throughput.avg.pct|issue_active.avg.pct|achieved_occupancy|warps_active.avg.pct|long_scoreboard_per|l1tex__t_sector_hit_rate|lts__t_sector_hit_rate.pct|registers_per_thread|launch__grid_size|lg_throttle_per|mio_throttle_per|l1tex__throughput.avg.pct|lts__throughput.avg.pct" gpurun_out/r2_ncu_full_tail_${NAME}.metrics.txt | head -20
done

#!/bin/bash
# Final measurement pass of round 2, second session (one gpurun call): bench lines (headline with cpu_baseline, reference arm,
# FPN/ResNet-50 --config c3, DeepLabV3 --config people + its reference arm), ncu launch lists of one eager step (headline, c3),
# `ncu --set full` extracts of the dominant conv (headline) and of the halo kernel serving a 1x1 layer as a plain GEMM (c3).
TAG=${1:-r2_s10}
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json.log 2>&1
tail -1 gpurun_out/${TAG}_bench.json.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json.log 2>&1
tail -1 gpurun_out/${TAG}_bench_reference.json.log | cut -c1-200
timeout 200 python bench.py --config c3 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c3.json.log 2>&1
tail -1 gpurun_out/${TAG}_bench_c3.json.log | cut -c1-200
timeout 300 python bench.py --config people --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_people.json.log 2>&1
tail -1 gpurun_out/${TAG}_bench_people.json.log | cut -c1-300
timeout 300 python bench.py --config people --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_people_reference.json.log 2>&1
tail -1 gpurun_out/${TAG}_bench_people_reference.json.log | cut -c1-200
for CFG in c2 c3 people; do
  LPS=$(timeout 300 python scripts/profile_step.py --config $CFG --steps 1 2>/dev/null | awk '/launches_per_step/{print $2}')
  echo "$CFG launches_per_step=$LPS" > gpurun_out/${TAG}_prof_${CFG}.log
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2*LPS)) -c $LPS --csv --log-file gpurun_out/${TAG}_launches_${CFG}.csv python scripts/profile_step.py --config $CFG --steps 3 >> gpurun_out/${TAG}_prof_${CFG}.log 2>&1
  python scripts/ncu_summary.py gpurun_out/${TAG}_launches_${CFG}.csv > gpurun_out/${TAG}_launches_${CFG}.summary.txt 2>/dev/null
  head -6 gpurun_out/${TAG}_launches_${CFG}.summary.txt | cut -c1-150
done
for spec in "tc3:c2:conv_tc3_kernel:20:2" "tc2_1x1:c3:conv_tc2_kernel:6:4"; do
  IFS=: read NAME CFG RE SKIP CNT <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$RE" -s $SKIP -c $CNT -f -o gpurun_out/${TAG}_ncu_${NAME} python scripts/profile_step.py --config $CFG --steps 1 > gpurun_out/${TAG}_ncu_${NAME}.log 2>&1
  ncu -i gpurun_out/${TAG}_ncu_${NAME}.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_${NAME}.raw.csv 2>/dev/null
  python scripts/ncu_raw.py gpurun_out/${TAG}_ncu_${NAME}.raw.csv > gpurun_out/${TAG}_ncu_full_${NAME}.metrics.txt 2>/dev/null
  rm -f gpurun_out/${TAG}_ncu_${NAME}.ncu-rep
  grep -E "^----|gpu__time_duration.sum|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|dram__bytes_read.sum |dram__bytes_write.sum " gpurun_out/${TAG}_ncu_full_${NAME}.metrics.txt | head -20 | cut -c1-160
done

#!/bin/bash
# Round-2 measurement pass (one gpurun call): headline bench (with cpu_baseline), the reference arm, the BN-backward-fusion A/B,
# the ncu launch list of one eager step, and `ncu --set full` captures of the dominant conv and the top streaming kernels.
TAG=${1:-r2_s4}
mkdir -p gpurun_out
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json.log 2>&1
tail -1 gpurun_out/${TAG}_bench.json.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json.log 2>&1
tail -1 gpurun_out/${TAG}_bench_reference.json.log | cut -c1-200
timeout 200 python bench.py --config c3 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_c3.json.log 2>&1
tail -1 gpurun_out/${TAG}_bench_c3.json.log | cut -c1-200
PYTHONPATH=. timeout 200 python scripts/deeplab_bench.py --top 40 > gpurun_out/${TAG}_deeplab_bench.txt 2>&1
head -1 gpurun_out/${TAG}_deeplab_bench.txt
timeout 300 python bench.py --config people --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_people.json.log 2>&1
tail -1 gpurun_out/${TAG}_bench_people.json.log | cut -c1-300
timeout 300 python bench.py --config people --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_people_reference.json.log 2>&1
tail -1 gpurun_out/${TAG}_bench_people_reference.json.log | cut -c1-200
LPS=$(timeout 300 python scripts/profile_step.py --steps 1 2>/dev/null | awk '/launches_per_step/{print $2}')
echo "launches_per_step=$LPS" > gpurun_out/${TAG}_prof.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s $((2*LPS)) -c $LPS --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/profile_step.py --steps 3 >> gpurun_out/${TAG}_prof.log 2>&1
tail -2 gpurun_out/${TAG}_prof.log
for spec in "tc3:conv_tc3_kernel:20:3" "bnred:reduce_rows_kernel:60:2" "bnbwd:bwd_apply_kernel:60:2" "bnapply:bn_apply_kernel:20:2"; do
  IFS=: read NAME RE SKIP CNT <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$RE" -s $SKIP -c $CNT -f -o gpurun_out/${TAG}_ncu_${NAME} python scripts/profile_step.py --steps 2 > gpurun_out/${TAG}_ncu_${NAME}.log 2>&1
  ncu -i gpurun_out/${TAG}_ncu_${NAME}.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_${NAME}.raw.csv 2>/dev/null
  python scripts/ncu_raw.py gpurun_out/${TAG}_ncu_${NAME}.raw.csv > gpurun_out/${TAG}_ncu_full_${NAME}.metrics.txt 2>/dev/null
  rm -f gpurun_out/${TAG}_ncu_${NAME}.ncu-rep
  grep -E "gpu__time_duration.sum|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|dram__bytes_read.sum|dram__bytes_write.sum" gpurun_out/${TAG}_ncu_full_${NAME}.metrics.txt | head -8
done

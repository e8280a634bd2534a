for N in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_s7_bench_${N}gpu.json.log 2>&1
tail -1 gpurun_out/r2_s7_bench_${N}gpu.json.log | cut -c1-230
done

mkdir -p gpurun_out
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_val_${N}gpu.json 2> gpurun_out/bench_val_${N}gpu.err; tail -2 gpurun_out/bench_val_${N}gpu.err; cat gpurun_out/bench_val_${N}gpu.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 10 --warmup 3 --workload fit1d --no-cpu --no-e2e > gpurun_out/bench_fit_8gpu.json 2> gpurun_out/bench_fit_8gpu.err; tail -2 gpurun_out/bench_fit_8gpu.err; cat gpurun_out/bench_fit_8gpu.json

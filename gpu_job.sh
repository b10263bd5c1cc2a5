mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_val_8gpu.json 2> gpurun_out/bench_val_8gpu.err; cut -c1-300 gpurun_out/bench_val_8gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload fit1d --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_fit_8gpu.json 2> gpurun_out/bench_fit_8gpu.err; cut -c1-300 gpurun_out/bench_fit_8gpu.json

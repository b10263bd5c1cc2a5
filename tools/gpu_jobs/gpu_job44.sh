mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py tests/test_gpu_threads.py -x -q 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu_val.json 2> gpurun_out/bench_2gpu.err; tail -2 gpurun_out/bench_2gpu.err; cut -c1-220 gpurun_out/r02_bench_2gpu_val.json

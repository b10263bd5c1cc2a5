mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_fit.py -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-legs > gpurun_out/r02_bench_2gpu_val.json 2> gpurun_out/bench_2gpu.err; tail -3 gpurun_out/bench_2gpu.err; cut -c1-300 gpurun_out/r02_bench_2gpu_val.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --mode jac --no-legs --no-cpu > gpurun_out/r02_bench_2gpu_jac.json 2> gpurun_out/bench_2gpu_jac.err; tail -3 gpurun_out/bench_2gpu_jac.err; cut -c1-300 gpurun_out/r02_bench_2gpu_jac.json

mkdir -p gpurun_out
N=${NGPU:-8}
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -3 gpurun_out/bench_${N}gpu.err; cut -c1-250 gpurun_out/r02_bench_${N}gpu.json
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3

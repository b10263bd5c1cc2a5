# 2-GPU record with the final library: value and Jacobian lines (fused peer-store gather), reference arm under torchrun, 2-GPU tests
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_2gpu_val.json 2> gpurun_out/r02_bench_2gpu_val.err; cut -c1-200 gpurun_out/r02_bench_2gpu_val.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --mode jac --no-legs > gpurun_out/r02_bench_2gpu_jac.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_2gpu_jac.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -3

set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_product.py tests/test_gpu_multi.py tests/test_gpu_eval.py tests/test_gpu_sharded.py -x -q 2>&1 | tail -15
QUICK=1 python bench_legs.py > gpurun_out/r02_legs_job3.json 2> gpurun_out/legs_job3.err; tail -3 gpurun_out/legs_job3.err
python bench.py --steps 5 --warmup 3 --no-legs --no-cpu > gpurun_out/r02_bench_job3.json 2> gpurun_out/bench_job3.err; tail -3 gpurun_out/bench_job3.err

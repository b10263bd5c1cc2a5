mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_product.py tests/test_gpu_eval.py tests/test_gpu_multi.py -x -q 2>&1 | tail -3
QUICK=1 python bench_legs.py > gpurun_out/r02_legs_job10.json 2> gpurun_out/legs_job10.err; tail -3 gpurun_out/legs_job10.err

mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_adrules.py tests/test_gpu_eval.py -m gpu -q -x 2>&1 | tail -30

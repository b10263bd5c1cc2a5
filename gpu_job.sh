timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_adrules.py tests/test_gpu_regression.py tests/test_gpu_roots.py -x -q -m gpu 2>&1 | tail -3

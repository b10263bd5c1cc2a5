timeout 1500 python -m pytest tests/test_gpu_fit.py tests/test_gpu_roots.py tests/test_gpu_golden.py tests/test_gpu_regression.py -x -q -m gpu 2>&1 | tail -3
python tools/latency_probe.py 2>&1 | tail -9

python -m pytest tests/test_gpu_fit.py tests/test_gpu_fit_nd.py tests/test_gpu_roots.py tests/test_gpu_golden.py tests/test_gpu_regression.py -x -q 2>&1 | tail -2
python tools/latency_probe.py 2>&1 | tee gpurun_out/r02_latency.txt | tail -15

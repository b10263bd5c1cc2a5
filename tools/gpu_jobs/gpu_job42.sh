python -m pytest tests/test_gpu_fit_nd.py tests/test_gpu_fit.py -x -q 2>&1 | tail -2
python tools/probe_fit_nd.py 2>&1 | cut -c1-45,95-175 | head -9

python -m pytest tests/test_gpu_loworder.py tests/test_gpu_guard_bands.py -x -q 2>&1 | tail -12

python -m pytest tests/test_gpu_guard_bands.py tests/test_gpu_product.py -x -q 2>&1 | tail -15

python -m pytest tests/test_gpu_loworder.py tests/test_gpu_guard_bands.py -x -q 2>&1 | tail -3
python tools/probe_loworder.py 2>&1 | grep -E "1d_order[48]_" | cut -c1-120,240-330

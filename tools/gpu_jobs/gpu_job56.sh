mkdir -p gpurun_out
python -m pytest tests/test_gpu_loworder.py -x -q 2>&1 | tail -2
python tools/probe_loworder.py 2>&1 | grep -E "order24|order64|order200|order512" | cut -c1-120,240-330

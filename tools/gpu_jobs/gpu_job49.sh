# low-order 1-D path: parity tests, probe, then the whole suite
mkdir -p gpurun_out
python -m pytest tests/test_gpu_loworder.py -x -q 2>&1 | tail -15
python tools/probe_loworder.py 2>&1 | tee gpurun_out/r02_loworder_v3.jsonl | cut -c1-120,240-330
python -m pytest tests -m gpu -x -q 2>&1 | tail -3

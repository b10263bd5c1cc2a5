timeout 300 python -m pytest tests/test_gpu_eval.py -x -q -k "shared or many" 2>&1 | tail -2
timeout 120 python tools/probe_shared.py 2>&1 | grep -v Trace

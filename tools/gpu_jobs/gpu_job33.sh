timeout 300 python -m pytest tests/test_gpu_eval.py -x -q -k "shared or many" 2>&1 | tail -2
echo RT1; timeout 120 python tools/probe_shared.py
echo RT2; FASTCHEB_SHARED_RT=2 timeout 120 python tools/probe_shared.py | head -3

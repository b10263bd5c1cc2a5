mkdir -p gpurun_out
python -m pytest tests/test_gpu_eval.py -x -q -k "shared or many" 2>&1 | tail -2
echo RT1; python tools/probe_shared.py
echo RT2; FASTCHEB_SHARED_RT=2 python tools/probe_shared.py
FASTCHEB_SHARED_RT=2 python -m pytest tests/test_gpu_eval.py -x -q -k "shared" 2>&1 | tail -2

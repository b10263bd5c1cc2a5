python -m pytest tests/test_gpu_eval.py -x -q -k "many" 2>&1 | tail -2
python tools/probe_many.py 2>&1 | grep -v Trace

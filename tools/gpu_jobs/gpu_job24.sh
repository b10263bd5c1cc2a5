python -m pytest tests/test_gpu_eval.py -x -q -k "many" 2>&1 | tail -3
echo REG; python tools/probe_many.py
echo REG P4; FASTCHEB_MANY_REG=4 python tools/probe_many.py | head -1
echo OLD; FASTCHEB_MANY_REG=0 python tools/probe_many.py

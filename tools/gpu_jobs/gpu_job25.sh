python -m pytest tests/test_gpu_eval.py -x -q -k "many" 2>&1 | tail -2
for m in 1 16 4 0; do echo "MANY_REG=$m"; FASTCHEB_MANY_REG=$m python tools/probe_many.py 2>&1 | grep -v Trace | sed -n 1,2p; done

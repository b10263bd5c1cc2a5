mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -5
BENCH_ONLY="cfg3" timeout 300 python tools/bench_configs.py 2>/dev/null | grep '"dmma' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'][:44], d['algo'], d['mode'], round(d['frac_of_fp_peak'], 4))"
BENCH_ONLY="cfg4" timeout 300 python tools/bench_configs.py 2>/dev/null | grep '"dmma' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config'][:44], d['algo'], d['mode'], round(d['frac_of_fp_peak'], 4))"

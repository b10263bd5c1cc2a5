timeout 900 python -m pytest tests/test_gpu_eval.py tests/test_gpu_adrules.py -x -q -m gpu 2>&1 | tail -3
export BENCH_ONLY="shape" BENCH_SHAPES="41x41,49x49,41x41x41"
timeout 600 python tools/bench_configs.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if 'shape' in d['config']: print(d['config'][:24].ljust(24), d['algo'].ljust(9), d['mode'], '%.3e pts/s' % d['points_per_s'], round(d['frac_of_fp_peak'], 3))"

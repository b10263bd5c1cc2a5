timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
export BENCH_SHAPES="5x8,9x9,9x9x9"
timeout 600 python tools/bench_configs.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d['algo']=='clenshaw' and ('shape' in d['config'] or 'cfg1' in d['config'] or 'cfg5' in d['config']): print(d['config'][:52].ljust(52), d['mode'], '%.3e pts/s' % d['points_per_s'], round(d['frac_of_fp_peak'], 3))"

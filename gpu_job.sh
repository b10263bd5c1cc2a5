mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eval.py -x -q -m gpu 2>&1 | tail -3
for st in 0 1000 2000 4000; do
FASTCHEB_2D_STAGGER=$st BENCH_ONLY="2-D" timeout 300 python tools/bench_configs.py 2>/dev/null | grep '"dmma"' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('stagger $st', d['config'][:34], d['mode'], round(d['frac_of_fp_peak'], 4))"
done

set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_roots.py tests/test_gpu_fit.py -m gpu -q -x 2>&1 | tail -30
python bench.py --workload fit1d --steps 20 --warmup 3 --dtype f32 --no-cpu > gpurun_out/bench_fit1d_f32.json; cat gpurun_out/bench_fit1d_f32.json
for tune in 24,3,1 24,4,1 20,4,1 16,6,1; do
  FASTCHEB_FIT_TUNE=$tune python bench.py --workload fit1d --steps 20 --warmup 3 --no-cpu --no-e2e --dtype f32 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print('TUNE-f32 $tune', d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['max_rel_err_vs_oracle'])
"
done

set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
python bench.py --workload fit1d --steps 20 --warmup 3 > gpurun_out/bench_fit1d_f64.json; cat gpurun_out/bench_fit1d_f64.json
python bench.py --workload fit1d --steps 20 --warmup 3 --dtype f32 > gpurun_out/bench_fit1d_f32.json; cat gpurun_out/bench_fit1d_f32.json
python bench.py --workload fit1d --steps 20 --warmup 3 --howmany 100000 --no-cpu > gpurun_out/bench_fit1d_f64_1e5.json; cat gpurun_out/bench_fit1d_f64_1e5.json
for tune in 16,3,1 24,2,1 16,2,1; do
  FASTCHEB_FIT_TUNE=$tune python bench.py --workload fit1d --steps 20 --warmup 3 --no-cpu --no-e2e --dtype f32 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print('TUNE-f32 $tune', d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['max_rel_err_vs_oracle'])
"
done
CMD="python bench.py --workload fit1d --steps 2 --warmup 3 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:fit_dmma_kernel -s 3 -c 1 -o gpurun_out/r01_fit_dmma_v4 -f $CMD > gpurun_out/ncu_fit2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r01_launches_fit.csv $CMD > gpurun_out/ncu_l.log 2>&1
tail -n 2 gpurun_out/ncu_fit2.log

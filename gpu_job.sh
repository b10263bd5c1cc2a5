set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit.py -m gpu -q -x 2>&1 | tail -5
for tune in 16,3,1 16,4,1 24,2,1 20,2,1 12,4,1 8,3,2; do
  FASTCHEB_FIT_TUNE=$tune python bench.py --workload fit1d --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print('TUNE $tune', d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['max_rel_err_vs_oracle'])
"
done
python bench.py --workload fit1d --steps 20 --warmup 3 --no-cpu --no-e2e --dtype f32 | cut -c1-900
CMD="python bench.py --workload fit1d --steps 2 --warmup 3 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:fit_dmma_kernel -s 3 -c 1 -o gpurun_out/r01_fit_dmma_v3 -f $CMD > gpurun_out/ncu_fit2.log 2>&1
tail -n 2 gpurun_out/ncu_fit2.log

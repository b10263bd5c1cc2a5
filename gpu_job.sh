mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fit.py tests/test_gpu_fullsize.py tests/test_gpu_sharded.py tests/test_gpu_threads.py -m gpu -q -x 2>&1 | tail -5
for mode in dyn static; do
  if [ $mode = static ]; then export FASTCHEB_FIT_STATIC=1; else unset FASTCHEB_FIT_STATIC; fi
  for hm in 1048576 100000 20000; do
    python bench.py --workload fit1d --steps 30 --warmup 5 --howmany $hm --no-cpu --no-e2e | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$mode', $hm, 'f64', round(d['value']/1e9,3), 'Gfits/s', round(d['roofline']['frac']*100,1), '%', d['max_rel_err_vs_oracle'])"
  done
  python bench.py --workload fit1d --steps 30 --warmup 5 --dtype f32 --no-cpu --no-e2e | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$mode', 1048576, 'f32', round(d['value']/1e9,3), 'Gfits/s', round(d['roofline']['frac']*100,1), '%', d['max_rel_err_vs_oracle'])"
done

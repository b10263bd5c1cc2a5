timeout 1500 python -m pytest tests/test_gpu_fit.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_roots.py tests/test_gpu_threads.py -x -q -m gpu 2>&1 | tail -3
for dt in f64 f32; do
python bench.py --workload fit1d --steps 20 --warmup 3 --dtype $dt --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$dt 2^20', round(d['roofline']['frac'],4), d['ms_per_step'], d['max_rel_err_vs_oracle'])"
python bench.py --workload fit1d --steps 50 --warmup 5 --dtype $dt --howmany 100000 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$dt 1e5', round(d['roofline']['frac'],4), d['ms_per_step'])"
python bench.py --workload fit1d --steps 50 --warmup 5 --dtype $dt --howmany 30000 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$dt 3e4', round(d['roofline']['frac'],4), d['ms_per_step'])"
done

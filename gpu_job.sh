mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fit.py tests/test_gpu_fullsize.py tests/test_gpu_sharded.py -m gpu -q -x 2>&1 | tail -8
python bench.py --workload fit1d --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_fit1d_f64.json; grep -o '"value": [0-9.]*\|"frac": [0-9.]*\|max_rel_err_vs_oracle": [0-9.e-]*' gpurun_out/bench_fit1d_f64.json
python bench.py --workload fit1d --steps 20 --warmup 3 --dtype f32 --no-cpu --no-e2e > gpurun_out/bench_fit1d_f32.json; grep -o '"value": [0-9.]*\|"frac": [0-9.]*\|max_rel_err_vs_oracle": [0-9.e-]*' gpurun_out/bench_fit1d_f32.json
python bench.py --workload fit1d --steps 20 --warmup 3 --howmany 100000 --no-cpu --no-e2e > gpurun_out/bench_fit1d_f64_1e5.json; grep -o '"value": [0-9.]*\|"frac": [0-9.]*' gpurun_out/bench_fit1d_f64_1e5.json

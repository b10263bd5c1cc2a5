set -x
mkdir -p gpurun_out
python bench.py --workload fit1d --steps 20 --warmup 3 > gpurun_out/bench_fit1d_f64.json
python bench.py --workload fit1d --steps 20 --warmup 3 --dtype f32 > gpurun_out/bench_fit1d_f32.json
python bench.py --workload fit1d --steps 50 --warmup 5 --howmany 100000 --no-cpu > gpurun_out/bench_fit1d_f64_1e5.json
python bench.py --workload fit1d --steps 50 --warmup 5 --howmany 100000 --no-cpu --dtype f32 > gpurun_out/bench_fit1d_f32_1e5.json
CMDF="python bench.py --workload fit1d --steps 2 --warmup 3 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r01_launches_fit.csv $CMDF > gpurun_out/ncu_lf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fit_dmma_kernel -s 3 -c 1 -o gpurun_out/r01_fit_dmma_v8 -f $CMDF > gpurun_out/ncu_ff.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fit_dmma_kernel -s 3 -c 1 -o gpurun_out/r01_fit_dmma_v8_f32 -f $CMDF --dtype f32 > gpurun_out/ncu_ff32.log 2>&1
grep -o '"frac": [0-9.]*' gpurun_out/bench_fit1d_*.json

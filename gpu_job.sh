set -x
mkdir -p gpurun_out
python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
python bench.py --workload fit1d --steps 20 --warmup 3 > gpurun_out/bench_fit1d_f64.json
python bench.py --workload fit1d --steps 20 --warmup 3 --dtype f32 > gpurun_out/bench_fit1d_f32.json
python bench.py --workload fit1d --steps 20 --warmup 3 --howmany 100000 --no-cpu > gpurun_out/bench_fit1d_f64_1e5.json
python bench.py --workload fit1d --steps 20 --warmup 3 --howmany 100000 --no-cpu --dtype f32 > gpurun_out/bench_fit1d_f32_1e5.json
CMDF="python bench.py --workload fit1d --steps 2 --warmup 3 --no-cpu --no-e2e --dtype f32"
ncu --set full --clock-control none --import-source on -k regex:fit_dmma_kernel -s 3 -c 1 -o gpurun_out/r01_fit_dmma_v7_f32 -f $CMDF > gpurun_out/ncu_ff32.log 2>&1
grep -o '"frac": [0-9.]*' gpurun_out/bench_fit1d_*.json

set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_val_p.json 2> gpurun_out/bench_val.err; tail -3 gpurun_out/bench_val.err; cat gpurun_out/bench_val_p.json
python bench.py --steps 5 --warmup 3 --mode jac --no-cpu --no-e2e > gpurun_out/bench_jac_p.json 2> gpurun_out/bench_jac.err; tail -3 gpurun_out/bench_jac.err; cat gpurun_out/bench_jac_p.json

set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_val_dmma.json 2> gpurun_out/bench_val_dmma.err; tail -3 gpurun_out/bench_val_dmma.err; cat gpurun_out/bench_val_dmma.json
python bench.py --steps 3 --warmup 3 --algo clenshaw --batch 4194304 --no-cpu --no-e2e > gpurun_out/bench_val_clenshaw.json 2> gpurun_out/bench_val_clenshaw.err; tail -3 gpurun_out/bench_val_clenshaw.err; cat gpurun_out/bench_val_clenshaw.json
python bench.py --steps 5 --warmup 3 --mode jac --no-cpu > gpurun_out/bench_jac_dmma.json 2> gpurun_out/bench_jac_dmma.err; tail -3 gpurun_out/bench_jac_dmma.err; cat gpurun_out/bench_jac_dmma.json
python bench.py --steps 3 --warmup 3 --mode jac --algo clenshaw --batch 4194304 --no-cpu --no-e2e > gpurun_out/bench_jac_clenshaw.json 2> gpurun_out/bench_jac_clenshaw.err; tail -3 gpurun_out/bench_jac_clenshaw.err; cat gpurun_out/bench_jac_clenshaw.json

# Standard verification job for a B200 box:  gpurun --timeout 2400 -- bash gpu_job.sh
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_default.json
python bench.py --workload fit1d --steps 20 --warmup 3 > gpurun_out/bench_fit1d_f64.json; grep -o '"frac": [0-9.]*' gpurun_out/bench_fit1d_f64.json

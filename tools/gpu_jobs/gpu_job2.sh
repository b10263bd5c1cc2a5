set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -5 gpurun_out/r02_bench_default.err; cut -c1-600 gpurun_out/r02_bench_default.json
python tools/latency_probe.py > gpurun_out/r02_latency.txt 2>&1; cat gpurun_out/r02_latency.txt

set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --impl reference > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json

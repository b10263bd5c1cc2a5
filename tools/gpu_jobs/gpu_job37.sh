mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err ) 2>&1 | grep real
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; cut -c1-160 gpurun_out/r02_bench_reference.json
python bench.py --mode jac --steps 10 --warmup 3 --no-legs > gpurun_out/r02_bench_jac.json 2>/dev/null; cut -c1-160 gpurun_out/r02_bench_jac.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_shared_kernel -s 1 -c 1 -f -o gpurun_out/r02_eval_shared_v5 python tools/probe_shared_once.py > /dev/null 2>&1
python tools/probe_fit_nd.py 2>/dev/null > gpurun_out/r02_fit_nd_probe.jsonl; wc -l gpurun_out/r02_fit_nd_probe.jsonl

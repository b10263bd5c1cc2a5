# full validation + artifact refresh with the low-order kernels; ncu summaries of the two low-order kernels; SASS histogram input
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err ) 2>&1 | grep real
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; cut -c1-160 gpurun_out/r02_bench_reference.json
python bench.py --mode jac --steps 10 --warmup 3 --no-legs > gpurun_out/r02_bench_jac.json 2>/dev/null; cut -c1-160 gpurun_out/r02_bench_jac.json
python tools/probe_loworder.py > gpurun_out/r02_loworder.jsonl 2>&1
FASTCHEB_LOND=0 FASTCHEB_LO1D=0 python tools/probe_loworder.py > gpurun_out/r02_loworder_generic.jsonl 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_lo1d_kernel -s 1 -c 1 -f -o gpurun_out/r02_lo1d_order8 python tools/probe_loworder.py once 2 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_lond_kernel -s 1 -c 1 -f -o gpurun_out/r02_lond_5x5 python tools/probe_loworder.py once 4 > /dev/null 2>&1
python tools/latency_probe.py > gpurun_out/r02_latency.txt 2>&1; tail -3 gpurun_out/r02_latency.txt
timeout 600 python tools/big_index_loworder.py 2>&1 | tail -4 | tee gpurun_out/r02_big_index_loworder.txt
timeout 300 python tools/leak_check.py 2>&1 | tail -2

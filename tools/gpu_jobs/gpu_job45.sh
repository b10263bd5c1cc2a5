# full validation + artifact refresh after the derivative-series gradient kernels; ncu capture of the order-200 gradient kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err ) 2>&1 | grep real
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; cut -c1-160 gpurun_out/r02_bench_reference.json
python bench.py --mode jac --steps 10 --warmup 3 --no-legs > gpurun_out/r02_bench_jac.json 2>/dev/null; cut -c1-160 gpurun_out/r02_bench_jac.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_pb1d_kernel -s 1 -c 1 -f -o gpurun_out/r02_pb1d_grad python tools/probe_pb1d.py once 201 grad > /dev/null 2>&1
ls -la gpurun_out/r02_pb1d_grad.ncu-rep

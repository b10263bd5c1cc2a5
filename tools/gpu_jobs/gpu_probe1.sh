set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
python tools/jac_parity.py > gpurun_out/r02_jac_parity.jsonl 2> gpurun_out/jac_parity.err; tail -3 gpurun_out/jac_parity.err; cat gpurun_out/r02_jac_parity.jsonl | cut -c1-400
# ncu --set full of the CUDA-core kernels that had no capture in round 1
NCU="ncu --set full --clock-control none --import-source on -c 1"
BENCH_ONLY="cfg1" timeout 600 $NCU -k regex:clenshaw_kernel -o gpurun_out/r02_clenshaw1d_cfg1 python tools/bench_configs.py > gpurun_out/ncu1.log 2>&1
BENCH_ONLY="cfg5 eval: 1-D order 64 real F64" timeout 600 $NCU -k regex:clenshaw_kernel -o gpurun_out/r02_clenshaw1d_cfg5 python tools/bench_configs.py > gpurun_out/ncu2.log 2>&1
timeout 600 $NCU -k regex:eval_many_kernel -o gpurun_out/r02_eval_many python tools/bench_configs.py many > gpurun_out/ncu3.log 2>&1
BENCH_ONLY="cfg3:" timeout 600 $NCU -k regex:clenshaw_kernel -o gpurun_out/r02_clenshaw3d_cfg3 python tools/bench_configs.py > gpurun_out/ncu4.log 2>&1
FIT_ONLY=65x65-float64 timeout 600 $NCU -k regex:dct1_direct_kernel -o gpurun_out/r02_dct1_direct python tools/probe_fit_nd.py > gpurun_out/ncu5.log 2>&1
python tools/probe_fit_nd.py > gpurun_out/fit_nd_baseline.jsonl 2>&1
ls -la gpurun_out

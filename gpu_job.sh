mkdir -p gpurun_out
BENCH_ONLY="cfg1" ncu --set full --clock-control none --import-source on -k regex:clenshaw_kernel -s 3 -c 1 -o gpurun_out/r01_clenshaw_cfg1 -f python tools/bench_configs.py > gpurun_out/ncu_s.log 2>&1
BENCH_ONLY="cfg5 eval: 1-D order 64 real F64" ncu --set full --clock-control none --import-source on -k regex:clenshaw_kernel -s 3 -c 1 -o gpurun_out/r01_clenshaw_cfg5 -f python tools/bench_configs.py > gpurun_out/ncu_s2.log 2>&1
tail -n 2 gpurun_out/ncu_s.log gpurun_out/ncu_s2.log

mkdir -p gpurun_out
export BENCH_ONLY="cfg2"
ncu --set full --clock-control none --import-source on -k regex:dmma2d_kernel -s 3 -c 1 -o gpurun_out/r01_dmma2d_val_v3 -f python tools/bench_configs.py > gpurun_out/ncu_2d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dmma2d_kernel -s 11 -c 1 -o gpurun_out/r01_dmma2d_jac_v3 -f python tools/bench_configs.py > gpurun_out/ncu_2dj.log 2>&1

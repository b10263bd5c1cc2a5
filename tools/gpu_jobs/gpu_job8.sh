mkdir -p gpurun_out
FIT_ONLY=65x65-float64 ncu --set full --clock-control none --import-source on -c 1 -s 3 -k regex:fit_nd_kernel -o gpurun_out/r02_fit_nd_spec python tools/probe_fit_nd.py > gpurun_out/ncu_fitnd.log 2>&1
tail -2 gpurun_out/ncu_fitnd.log

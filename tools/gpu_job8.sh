mkdir -p gpurun_out
FASTCHEB_FITND_WARPS=4 FASTCHEB_FITND_NBUF=1 FIT_ONLY=65x65-float64 ncu --set full --clock-control none --import-source on -c 1 -s 3 -k regex:fit_nd_kernel -o gpurun_out/r02_fit_nd_65x65 python tools/probe_fit_nd.py > gpurun_out/ncu_fitnd.log 2>&1
tail -2 gpurun_out/ncu_fitnd.log

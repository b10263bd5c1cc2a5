mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_lo1d_kernel -s 1 -c 1 -f -o gpurun_out/r02_lo1d_order64 python tools/probe_loworder.py once 20 > /dev/null 2>&1
ls -la gpurun_out/r02_lo1d_order64.ncu-rep

mkdir -p gpurun_out
for i in 0 1 3; do timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_lo1d_kernel -s 1 -c 1 -f -o gpurun_out/r02_lo1d_$i python tools/probe_loworder.py once $i > /dev/null 2>&1; done
ls -la gpurun_out/r02_lo1d_*.ncu-rep

mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_many_kernel -s 1 -c 1 -f -o gpurun_out/r02_eval_many_v2 python tools/probe_many_once.py > /dev/null 2>&1
ls -la gpurun_out/r02_eval_many_v2*

timeout 300 python -m pytest tests/test_gpu_eval.py -x -q -k "shared or many" 2>&1 | tail -2
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_shared_kernel -s 1 -c 1 -f -o gpurun_out/r02_eval_shared_v3 python tools/probe_shared_once.py > /dev/null 2>&1
ls -la gpurun_out/r02_eval_shared_v3*

timeout 600 python -m pytest tests/test_gpu_eval.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -2
timeout 120 python tools/probe_shared.py 2>&1 | grep -v Trace
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_shared_kernel -s 1 -c 1 -f -o gpurun_out/r02_eval_shared_v4 python tools/probe_shared_once.py > /dev/null 2>&1

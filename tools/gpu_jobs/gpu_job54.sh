# eval_many with consecutive points per lane / vector I/O / per-polynomial division test: parity + probe
mkdir -p gpurun_out
python -m pytest tests/test_gpu_eval.py tests/test_gpu_fuzz.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3
python tools/probe_many.py 2>&1 | tail -6 | tee gpurun_out/r02_many_probe_job54.txt

mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fit_nd.py tests/test_gpu_fit.py -x -q 2>&1 | tail -3
for w in "" 4 5 8; do echo "SPEC warps=$w"; FASTCHEB_FITND_WARPS=$w FIT_ONLY=65x65 timeout 120 python tools/probe_fit_nd.py 2>&1 | cut -c1-45,95-175; done
FIT_ONLY=33x33x33-float32 timeout 120 python tools/probe_fit_nd.py 2>&1 | cut -c1-45,95-175

mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit_nd.py tests/test_gpu_fit.py -x -q 2>&1 | tail -5
for w in "" 4 8; do for nb in 2 1; do echo "warps=$w nbuf=$nb"; FASTCHEB_FITND_WARPS=$w FASTCHEB_FITND_NBUF=$nb FIT_ONLY=65x65 python tools/probe_fit_nd.py 2>&1 | cut -c1-30,95-175; done; done
FIT_ONLY=33x33x33-float32 python tools/probe_fit_nd.py 2>&1 | cut -c1-175
FIT_ONLY=129 python tools/probe_fit_nd.py 2>&1 | cut -c1-175
FIT_ONLY=257 python tools/probe_fit_nd.py 2>&1 | cut -c1-175
FIT_ONLY=201 python tools/probe_fit_nd.py 2>&1 | cut -c1-175

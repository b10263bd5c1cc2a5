mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/latency_probe.py 2>&1 | tee gpurun_out/r02_latency.txt
FIT_ONLY=65x65-float64 python tools/probe_fit_nd.py 2>&1 | cut -c1-45,95-175

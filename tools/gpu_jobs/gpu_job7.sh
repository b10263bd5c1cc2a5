mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fit_nd.py tests/test_gpu_fit.py tests/test_gpu_product.py -x -q 2>&1 | tail -15
python tools/probe_fit_nd.py 2>&1 | tee gpurun_out/r02_fit_nd_probe.jsonl | cut -c1-200
python tools/latency_probe.py 2>&1 | head -8
python /tmp/pb.py 2>/dev/null || true

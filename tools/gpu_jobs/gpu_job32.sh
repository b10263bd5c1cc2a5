python -m pytest tests/test_gpu_fit.py -x -q 2>&1 | tail -2
for a in 0 12 10 8; do echo "AREG=$a"; for s in 65-float32; do FASTCHEB_FIT_AREG=$a FIT_ONLY=$s python tools/probe_fit_nd.py 2>&1 | cut -c1-45,95-175; done; done
echo "AREG=12 S=4"; FASTCHEB_FIT_TUNE=12,4,1,1 FIT_ONLY=65-float32 python tools/probe_fit_nd.py 2>&1 | cut -c1-45,95-175
echo "AREG=12 S=5 lag2"; FASTCHEB_FIT_TUNE=12,5,1,2 FIT_ONLY=65-float32 python tools/probe_fit_nd.py 2>&1 | cut -c1-45,95-175

python -m pytest tests/test_gpu_fit_nd.py tests/test_gpu_fit.py -x -q 2>&1 | tail -2
echo BLOCKS; python tools/probe_fit_nd.py 2>&1 | cut -c1-50,95-175 | head -10
echo NOBLOCKS; FASTCHEB_FITND_BLOCK=1 python tools/probe_fit_nd.py 2>&1 | cut -c1-50,95-175 | head -10

mkdir -p gpurun_out
python -m pytest tests/test_gpu_fit_nd.py tests/test_gpu_fit.py -x -q 2>&1 | tail -3
echo SPLIT; python tools/probe_fit_nd.py 2>&1 | cut -c1-45,95-200
echo NOSPLIT; FASTCHEB_FITND_NOSPLIT=1 python tools/probe_fit_nd.py 2>&1 | cut -c1-45,95-200 | head -6

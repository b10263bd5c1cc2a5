# product-basis gradient from the derivative series (second table), 2 points per thread for scalar gradients
mkdir -p gpurun_out
python -m pytest tests/test_gpu_product.py tests/test_gpu_adrules.py tests/test_gpu_jac_parity.py -x -q 2>&1 | tail -5
python tools/probe_pb1d.py 2>&1 | tail -8 | tee gpurun_out/r02_pb1d_probe_job44.txt
STRESS_SHARED=0 STRESS_FITS=0 STRESS_ND=0 python tools/stress_r02.py 2>&1 | tail -6 | tee gpurun_out/r02_stress_job44.txt
python -m pytest tests -m gpu -x -q 2>&1 | tail -3

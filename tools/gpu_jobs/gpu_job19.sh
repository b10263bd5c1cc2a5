mkdir -p gpurun_out
python -m pytest tests/test_gpu_product.py -x -q 2>&1 | tail -2
echo default; python tools/probe_pb1d.py 2>&1 | grep cfg
echo smallP; FASTCHEB_PB_SMALLP=1 python tools/probe_pb1d.py 2>&1 | grep f32
for t in 128 192; do echo threads $t; FASTCHEB_PB_THREADS=$t python tools/probe_pb1d.py 2>&1 | grep cfg; done

mkdir -p gpurun_out
python -m pytest tests/test_gpu_product.py -x -q 2>&1 | tail -3
python tools/probe_pb1d.py 2>&1 | tail -8
FASTCHEB_PB_A=16 python tools/probe_pb1d.py 2>&1 | grep cfg4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_pb1d -s 2 -c 1 -f -o gpurun_out/r02_pb1d_o64 python tools/probe_pb1d.py once 65 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_pb1d -s 2 -c 1 -f -o gpurun_out/r02_pb1d_o200b python tools/probe_pb1d.py once 201 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3

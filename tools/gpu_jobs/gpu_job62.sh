# refresh the ncu summaries of the restructured product-basis (order 64, order 200) and eval_many kernels
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_pb1d_kernel -s 1 -c 1 -f -o gpurun_out/r02_pb1d_o64_final python tools/probe_pb1d.py once 65 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_pb1d_kernel -s 1 -c 1 -f -o gpurun_out/r02_pb1d_o200_final python tools/probe_pb1d.py once 201 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_many_kernel -s 1 -c 1 -f -o gpurun_out/r02_eval_many_final python tools/probe_many_once.py > /dev/null 2>&1
ls -la gpurun_out/*_final.ncu-rep

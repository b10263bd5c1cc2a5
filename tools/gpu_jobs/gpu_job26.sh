mkdir -p gpurun_out
FASTCHEB_MANY_REG=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_many_reg -s 1 -c 1 -f -o gpurun_out/r02_many_reg_p4 python tools/probe_many_once.py > /dev/null 2>&1
FASTCHEB_MANY_REG=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:eval_many_reg -s 1 -c 1 -f -o gpurun_out/r02_many_reg_p8 python tools/probe_many_once.py > /dev/null 2>&1
ls -la gpurun_out/r02_many_reg*

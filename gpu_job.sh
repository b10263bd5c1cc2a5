set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dmma_kernel -s 3 -c 1 -o gpurun_out/r01_dmma_val_v7 -f $CMD > gpurun_out/ncu_f.log 2>&1
CMDJ="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --mode jac --batch 4194304"
ncu --set full --clock-control none --import-source on -k regex:dmma_kernel -s 3 -c 1 -o gpurun_out/r01_dmma_jac_v7 -f $CMDJ > gpurun_out/ncu_fj.log 2>&1
CMDF="python bench.py --workload fit1d --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMDF > gpurun_out/ncu_plain_f.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r01_launches_fit.csv $CMDF > gpurun_out/ncu_lf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fit_dmma_kernel -s 3 -c 1 -o gpurun_out/r01_fit_dmma_v5 -f $CMDF > gpurun_out/ncu_ff.log 2>&1
python bench.py --steps 5 --warmup 3 --mode jac --no-cpu > gpurun_out/bench_jac.json 2>/dev/null; cat gpurun_out/bench_jac.json | cut -c1-200
python bench.py --workload fit1d --steps 20 --warmup 3 > gpurun_out/bench_fit1d_f64.json; python bench.py --workload fit1d --steps 20 --warmup 3 --dtype f32 > gpurun_out/bench_fit1d_f32.json; python bench.py --workload fit1d --steps 20 --warmup 3 --howmany 100000 --no-cpu > gpurun_out/bench_fit1d_f64_1e5.json
grep -o '"frac": [0-9.]*' gpurun_out/bench_fit1d_f64.json gpurun_out/bench_fit1d_f32.json gpurun_out/bench_fit1d_f64_1e5.json

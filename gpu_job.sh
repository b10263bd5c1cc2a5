set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_default.json
python bench.py --steps 5 --warmup 3 --mode jac --no-cpu > gpurun_out/bench_jac.json 2>/dev/null; cut -c1-200 gpurun_out/bench_jac.json
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
CMDF="python bench.py --workload fit1d --steps 2 --warmup 3 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r01_launches_fit.csv $CMDF > gpurun_out/ncu_lf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fit_dmma_kernel -s 3 -c 1 -o gpurun_out/r01_fit_dmma_v7 -f $CMDF > gpurun_out/ncu_ff.log 2>&1

set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_val.json 2> gpurun_out/bench_val.err; tail -3 gpurun_out/bench_val.err; cat gpurun_out/bench_val.json
python bench.py --steps 5 --warmup 3 --mode jac --no-cpu > gpurun_out/bench_jac.json 2> gpurun_out/bench_jac.err; tail -3 gpurun_out/bench_jac.err; cat gpurun_out/bench_jac.json
CMD="python bench.py --steps 2 --warmup 3 --batch 1048576 --no-cpu --no-e2e"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dmma_kernel -s 3 -c 1 -o gpurun_out/r01_dmma_val_v4 -f $CMD > gpurun_out/ncu_f.log 2>&1
tail -n 2 gpurun_out/ncu_f.log

set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_val.json 2> gpurun_out/bench_val.err; tail -3 gpurun_out/bench_val.err; cat gpurun_out/bench_val.json
python bench.py --steps 5 --warmup 3 --mode jac --no-cpu > gpurun_out/bench_jac.json 2> gpurun_out/bench_jac.err; tail -3 gpurun_out/bench_jac.err; cat gpurun_out/bench_jac.json
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-500
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dmma_kernel -s 3 -c 1 -o gpurun_out/r01_dmma_val_v6 -f $CMD > gpurun_out/ncu_f.log 2>&1
tail -n 2 gpurun_out/ncu_f.log
CMDJ="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --mode jac --batch 4194304"
ncu --set full --clock-control none --import-source on -k regex:dmma_kernel -s 3 -c 1 -o gpurun_out/r01_dmma_jac_v6 -f $CMDJ > gpurun_out/ncu_fj.log 2>&1
tail -n 2 gpurun_out/ncu_fj.log
python -c "import __graft_entry__ as g; g.smoke()"

mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_eval.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_adrules.py -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_val.json 2> gpurun_out/bench_val.err; tail -3 gpurun_out/bench_val.err; grep -o '"value": [0-9.]*\|"frac": [0-9.]*' gpurun_out/bench_val.json
python bench.py --steps 5 --warmup 3 --mode jac --no-cpu --no-e2e > gpurun_out/bench_jac.json 2> gpurun_out/bench_jac.err; tail -3 gpurun_out/bench_jac.err; grep -o '"value": [0-9.]*\|"frac": [0-9.]*' gpurun_out/bench_jac.json
python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; tail -3 gpurun_out/configs.err
python - <<'PY'
import json
for l in open('gpurun_out/configs.jsonl'):
    d=json.loads(l)
    if d['algo']=='dmma': print(f"{d['config']:48s} {d['algo']:8s} {d['mode']} {d['points_per_s']:.3e} pts/s  fp {d['frac_of_fp_peak']*100:5.1f}%")
PY

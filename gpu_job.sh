mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_eval.py tests/test_gpu_golden.py tests/test_gpu_adrules.py -m gpu -q -x 2>&1 | tail -4
BENCH_ONLY="1-D" python tools/bench_configs.py > gpurun_out/configs1d.jsonl 2> gpurun_out/configs.err; tail -3 gpurun_out/configs.err
BENCH_ONLY="cfg1" python tools/bench_configs.py >> gpurun_out/configs1d.jsonl 2> gpurun_out/configs.err; tail -3 gpurun_out/configs.err
python - <<'PY'
import json
for l in open('gpurun_out/configs1d.jsonl'):
    d=json.loads(l)
    print(f"{d['config']:60s} {d['algo']:8s} {d['mode']} {d['points_per_s']:.3e} pts/s  fp {d['frac_of_fp_peak']*100:5.1f}%  hbm {d['frac_of_hbm']*100:5.1f}%")
PY

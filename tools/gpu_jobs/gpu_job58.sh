mkdir -p gpurun_out
python -m pytest tests/test_gpu_loworder.py tests/test_gpu_guard_bands.py -x -q 2>&1 | tail -4
python tools/probe_loworder.py 2>&1 > gpurun_out/r02_loworder_v6.jsonl
FASTCHEB_LO1D=0 FASTCHEB_LOND=0 python tools/probe_loworder.py 2>&1 > gpurun_out/r02_loworder_v6_off.jsonl
python - <<'PY'
import json
a=[json.loads(l) for l in open('gpurun_out/r02_loworder_v6.jsonl') if l.startswith('{')]
b=[json.loads(l) for l in open('gpurun_out/r02_loworder_v6_off.jsonl') if l.startswith('{')]
for x,y in zip(a,b):
    print(f"{x['leg']:28s} lo {x['points_per_s']:.3e}  generic {y['points_per_s']:.3e}  x{x['points_per_s']/y['points_per_s']:.2f}  hbm {x['frac_hbm']:.3f} fp {x['frac_fp']:.3f} {x['algo']}")
PY

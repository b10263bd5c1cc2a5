# lo1d as the 1-D Clenshaw kernel up to 513 coefficients: parity, probe with the path on / off
mkdir -p gpurun_out
python -m pytest tests/test_gpu_loworder.py tests/test_gpu_product.py -x -q 2>&1 | tail -5
python tools/probe_loworder.py 2>&1 > gpurun_out/r02_loworder_v5.jsonl
FASTCHEB_LO1D=0 FASTCHEB_LOND=0 python tools/probe_loworder.py 2>&1 > gpurun_out/r02_loworder_v5_off.jsonl
python - <<'PY'
import json
a=[json.loads(l) for l in open('gpurun_out/r02_loworder_v5.jsonl') if l.startswith('{')]
b=[json.loads(l) for l in open('gpurun_out/r02_loworder_v5_off.jsonl') if l.startswith('{')]
for x,y in zip(a,b):
    print(f"{x['leg']:28s} lo {x['points_per_s']:.3e}  generic {y['points_per_s']:.3e}  x{x['points_per_s']/y['points_per_s']:.2f}  hbm {x['frac_hbm']:.3f} fp {x['frac_fp']:.3f}")
PY
python -m pytest tests -m gpu -x -q 2>&1 | tail -3

mkdir -p gpurun_out
python tools/probe_crossover.py > gpurun_out/cross_auto.jsonl 2>&1
FASTCHEB_LOND_FIRST=1 python tools/probe_crossover.py > gpurun_out/cross_lond.jsonl 2>&1
python - <<'PY'
import json
a=[json.loads(l) for l in open('gpurun_out/cross_auto.jsonl') if l.startswith('{')]
b=[json.loads(l) for l in open('gpurun_out/cross_lond.jsonl') if l.startswith('{')]
for x,y in zip(a,b):
    if 'error' in x or 'error' in y: print(x, y); continue
    print(f"{str(x['dims']):16s} K={x['K']} jac={x['jac']!s:5s} auto[{x['algo']}] {x['points_per_s']:.3e}  lond-first {y['points_per_s']:.3e}  x{y['points_per_s']/x['points_per_s']:.2f}")
PY

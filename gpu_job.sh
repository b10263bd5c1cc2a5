mkdir -p gpurun_out
python tools/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --steps 5 --warmup 3 --mode jac --no-cpu > gpurun_out/bench_jac.json 2>/dev/null
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python -c "
import json
for f in ['bench_default','bench_jac','bench_ref']:
    d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][0]); print(f, d['value'], d.get('roofline',{}).get('frac'), d.get('e2e',{}).get('value'))"

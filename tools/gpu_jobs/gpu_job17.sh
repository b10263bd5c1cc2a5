mkdir -p gpurun_out
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, fastcheb
xg = fastcheb.chebpoints(200, 0.0, 10.0)
c = fastcheb.chebinterp(np.sin(2 * xg + 3 * np.cos(4 * xg)), 0.0, 10.0, tol=0.0)
print("product_check README order 200:", c.product_check())
x64 = fastcheb.chebpoints(64, -1.0, 1.0)
c2 = fastcheb.chebinterp(np.sin(2.3 * x64 + 0.4), -1.0, 1.0, tol=0.0)
print("product_check sin order 64:", c2.product_check())
rng = np.random.default_rng(0)
c3 = fastcheb.ChebPoly(rng.uniform(-1, 1, 201), [0.0], [10.0])
print("product_check dense 201:", c3.product_check())
PY
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; tail -2 gpurun_out/r02_bench_default.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_reference.json
python bench.py --mode jac --steps 10 --warmup 3 --no-legs > gpurun_out/r02_bench_jac.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_jac.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_val.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-legs --no-parity > /dev/null 2>&1
tail -5 gpurun_out/r02_launches_val.csv

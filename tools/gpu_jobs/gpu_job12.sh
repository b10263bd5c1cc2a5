mkdir -p gpurun_out
cat > /tmp/sh.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench_legs
from fastcheb import _capi
r = bench_legs.eval_many_leg(torch, _capi, 0, np.float64, False, 3, shared=True)
print("%.4g pts/s frac %.3f" % (r["value"], r["roofline"]["frac"]))
PY
ncu --set full --clock-control none --import-source on -c 1 -s 2 -k regex:eval_shared_kernel -o gpurun_out/r02_eval_shared python /tmp/sh.py > gpurun_out/ncu_sh.log 2>&1
tail -2 gpurun_out/ncu_sh.log

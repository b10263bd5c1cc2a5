mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eval.py -x -q -k "many" 2>&1 | tail -12
for rt in 1 2; do export FASTCHEB_SHARED_RT=$rt; echo RT=$rt; python - <<"PY"
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench_legs
from fastcheb import _capi
for dt in (np.float64, np.float32):
    for der in (False, True):
        r = bench_legs.eval_many_leg(torch, _capi, 0, dt, der, 5, shared=True)
        print(np.dtype(dt).name, "der" if der else "val", "%.4g pts/s  %.3f ms  frac %.3f hbm %.3f" % (r["value"], r["ms"], r["roofline"]["frac"], r["roofline"]["frac_hbm"]))
PY
done

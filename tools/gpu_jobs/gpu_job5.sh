mkdir -p gpurun_out
cat > /tmp/pb.py <<'PY'
import sys, json, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, fastcheb, bench_legs
from fastcheb import _capi
xg = fastcheb.chebpoints(200, 0.0, 10.0)
c1 = np.asarray(fastcheb.chebinterp(np.sin(2 * xg + 3 * np.cos(4 * xg)), 0.0, 10.0, tol=0.0).coefs)
for jac in (False, True):
    r = bench_legs.eval_leg(torch, fastcheb, _capi, 0, (201,), None, False, np.float64, 1 << 24, jac, 5, coefs=c1, lb=[0.0], ub=[10.0])
    print("jac" if jac else "val", r["algo"], "%.4g pts/s frac %.3f" % (r["value"], r["roofline"]["frac"]))
PY
python /tmp/pb.py
ncu --set full --clock-control none --import-source on -c 1 -s 2 -k regex:eval_pb1d_kernel -o gpurun_out/r02_pb1d_val python /tmp/pb.py > gpurun_out/ncu_pb.log 2>&1
tail -3 gpurun_out/ncu_pb.log

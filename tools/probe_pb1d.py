#!/usr/bin/env python
"""1-D evaluation legs only (configs[0] order 200, configs[4] order 64) + the product-basis self-check ratios.
usage: probe_pb1d.py [once N [grad]]   ("once": a single order-N launch, for ncu)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastcheb
from fastcheb import _capi
import bench_legs as B

f64, f32 = np.float64, np.float32
xg = fastcheb.chebpoints(200, 0.0, 10.0)
c1 = np.asarray(fastcheb.chebinterp(np.sin(2 * xg + 3 * np.cos(4 * xg)), 0.0, 10.0, tol=0.0).coefs)
x64 = fastcheb.chebpoints(64, -1.0, 1.0)
c64 = np.asarray(fastcheb.chebinterp(np.sin(2.3 * x64 + 0.4), -1.0, 1.0, tol=0.0).coefs)
if len(sys.argv) > 2 and sys.argv[1] == "once":
    n = int(sys.argv[2])
    co, lb, ub = (c1, [0.0], [10.0]) if n == 201 else (c64, [-1.0], [1.0])
    r = B.eval_leg(torch, fastcheb, _capi, 0, (n,), None, False, f64, 1 << 24, len(sys.argv) > 3 and sys.argv[3] == "grad", 1, coefs=co, lb=lb, ub=ub)
    print(json.dumps(r))
    sys.exit(0)
reps = 5
out = {}
for jac in (False, True):
    out["cfg0_" + ("grad" if jac else "val")] = B.eval_leg(torch, fastcheb, _capi, 0, (201,), None, False, f64, 1 << 24, jac, reps, coefs=c1, lb=[0.0], ub=[10.0])
    out["cfg4_f64_" + ("grad" if jac else "val")] = B.eval_leg(torch, fastcheb, _capi, 0, (65,), None, False, f64, 1 << 25, jac, reps, coefs=c64)
out["cfg4_f32_val"] = B.eval_leg(torch, fastcheb, _capi, 0, (65,), None, False, f32, 1 << 25, False, reps, coefs=c64.astype(f32))
for k, v in out.items():
    print(k, v["algo"], f"{v['value']:.4g} points/s  frac {v['roofline']['frac']:.4f}")
print("product_check README 200:", fastcheb.ChebPoly(c1, [0.0], [10.0]).product_check())
print("product_check sin 64:", fastcheb.ChebPoly(c64, [-1.0], [1.0]).product_check())

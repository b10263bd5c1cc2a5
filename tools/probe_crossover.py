#!/usr/bin/env python
"""Where the tensor-core path and the low-order nest cross (run twice: plain = AUTO, and with FASTCHEB_LOND_FIRST=1 =
low-order nest wherever it applies).  One line per shape: points/s."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastcheb
from fastcheb import _capi
import bench_legs as B
f64 = np.float64
SHAPES = [((9, 9), None), ((13, 13), None), ((17, 17), None), ((17, 33), None), ((33, 33), None), ((9, 40), None), ((13, 60), None),
          ((9, 9, 9), None), ((10, 10, 10), None), ((12, 12, 12), None), ((6, 6, 6, 6), None), ((17, 17), 3), ((9, 9, 9), 2)]
for dims, K in SHAPES:
    for jac in (False, True):
        try:
            r = B.eval_leg(torch, fastcheb, _capi, 0, dims, K, False, f64, 1 << 22, jac, 3)
            print(json.dumps({"dims": dims, "K": K, "jac": jac, "algo": r["algo"], "points_per_s": r["value"]}))
        except Exception as e:
            print(json.dumps({"dims": dims, "K": K, "jac": jac, "error": str(e)[:100]}))

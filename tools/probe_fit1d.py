#!/usr/bin/env python
"""Order-64 batched fit legs only (configs[4]): 10^5 and 2^20 fits, both precisions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastcheb
from fastcheb import _capi
import bench_legs as B
for dt in (np.float64, np.float32):
    for hm, reps in ((100_000, 40), (1 << 20, 20)):
        r = B.fit_leg(torch, _capi, 0, (65,), hm, dt, reps)
        print(np.dtype(dt).name, hm, f"{r['value']:.4g} fits/s  hbm {r['roofline']['frac']:.3f}")

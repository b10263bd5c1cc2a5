#!/usr/bin/env python
"""Per-polynomial-points eval_many legs only (configs[4]: 1e5 order-64 polynomials x 1000 own points)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastcheb
from fastcheb import _capi
import bench_legs as B
for dt in (np.float64, np.float32):
    for der in (False, True):
        r = B.eval_many_leg(torch, _capi, 0, dt, der, 5, shared=False)
        print("many", np.dtype(dt).name, "der" if der else "val", f"{r['value']:.4g} points/s frac {r['roofline']['frac']:.4f}")
for n in (17, 33):
    r = B.eval_many_leg(torch, _capi, 0, np.float64, False, 5, shared=False, n=n)
    print("many f64 n", n, f"{r['value']:.4g} points/s frac {r['roofline']['frac']:.4f}")

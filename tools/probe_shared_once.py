import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fastcheb import _capi
import bench_legs as B
r = B.eval_many_leg(torch, _capi, 0, np.float64, False, 1, shared=True)
print(r["value"], r["roofline"]["frac"])

#!/usr/bin/env python
"""Create / evaluate / destroy handles and run fits in a loop; device memory in use must not grow.
    python tools/leak_check.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastcheb  # noqa: E402

rng = np.random.default_rng(0)
c3 = rng.uniform(-1, 1, (17, 17, 17, 3))
c2 = rng.uniform(-1, 1, (65, 65))
x3 = rng.uniform(-1, 1, (1000, 3))
x2 = rng.uniform(-1, 1, (1000, 2))
vals = rng.uniform(-1, 1, (4096, 65))


def used():
    torch.cuda.synchronize()
    free, total = torch.cuda.mem_get_info()
    return (total - free) / 2**20


def cycle():
    p = fastcheb.ChebPoly(c3, -np.ones(3), np.ones(3)); p(x3); fastcheb.chebjacobian(p, x3); p.close()
    q = fastcheb.ChebPoly(c2, -np.ones(2), np.ones(2)); q(x2); fastcheb.chebgradient(q, x2); q.close()
    fastcheb.chebcoefs(vals, ndim=1, howmany=len(vals))
    fastcheb.chebinterp(c2, [0, 0], [1, 1]).close()


for _ in range(20):
    cycle()
m0 = used()
for _ in range(500):
    cycle()
m1 = used()
print(f"device memory in use: {m0:.1f} MiB after warm-up, {m1:.1f} MiB after 500 more cycles")
assert m1 - m0 < 8.0, "device memory grew"
print("leak check ok")

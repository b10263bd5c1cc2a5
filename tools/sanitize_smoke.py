#!/usr/bin/env python
"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Sizes are tiny: the sanitizer slows kernels down by 10-100x."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastcheb  # noqa: E402

rng = np.random.default_rng(0)
# evaluation: tensor-core and Clenshaw kernels, value and Jacobian, resident and slab-tiled tensors
for dims, K, cplx in (((13, 11, 9), 3, False), ((17, 17, 17, 17), None, True), ((65, 65), None, False), ((201,), None, False)):
    shape = dims + ((K,) if K else ())
    c = rng.uniform(-1, 1, shape)
    if cplx:
        c = c + 1j * rng.uniform(-1, 1, shape)
    N = len(dims)
    p = fastcheb.ChebPoly(c, -np.ones(N), np.ones(N))
    pts = rng.uniform(-1, 1, (3000, N))
    for algo in (fastcheb.ALGO_DMMA, fastcheb.ALGO_CLENSHAW):
        try:
            p.set_algo(algo)
        except Exception:
            continue
        p(pts)
        fastcheb.chebjacobian(p, pts)
    p.close()
# fits: tensor-core batched (one and two folds, ragged tail, K-vector, float32), generic N-d, droptol, interp
for n in (65, 40, 17):
    fastcheb.chebcoefs(rng.uniform(-1, 1, (1003, n)), ndim=1, howmany=1003, return_maxabs=True)
fastcheb.chebcoefs(rng.uniform(-1, 1, (300, 33, 3)).astype(np.float32), ndim=1, howmany=300)
bad = np.ones((600, 65))
bad[17, 3] = np.nan
try:
    fastcheb.chebcoefs(bad, ndim=1, howmany=600)
except fastcheb.DomainError:
    pass
fastcheb.chebinterp(rng.uniform(-1, 1, (20, 15, 4)), [0, 0, 0], [1, 1, 1])
fastcheb.eval_many(rng.uniform(-1, 1, (50, 65)), -1.0, 1.0, rng.uniform(-1, 1, (50, 200)), deriv=True)
c = fastcheb.chebinterp(np.cos, 200, 0, 30)
assert len(fastcheb.roots(c)) == 10
print("sanitize smoke done")

#!/usr/bin/env python
"""One-off 64-bit-index check of the low-order kernels: 2^31 + 5 device-resident points of a 1-D order-4 Float64 polynomial in
ONE call (17 GB in, 17 GB out), and 2^30 + 3 points of a 5 x 5 polynomial; the tails beyond 2^31 elements are compared
bitwise with a separate call on the same points, and both ends with the CPU oracle.  An out-of-domain point planted past
index 2^31 must be reported with its exact index."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import cheb_oracle as orc
import fastcheb

rng = np.random.default_rng(12)
for dims, n in (((5,), (1 << 31) + 5), ((5, 5), (1 << 30) + 3)):
    N = len(dims)
    c = rng.uniform(-1, 1, dims)
    lb, ub = -np.ones(N), np.ones(N)
    p = fastcheb.ChebPoly(c, lb, ub)
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    x = torch.empty((n, N), dtype=torch.float64, device="cuda")
    step = 1 << 27
    for i in range(0, n, step):
        x[i:i + step] = torch.rand((min(step, n - i), N), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    xx = x if N > 1 else x[:, 0]
    v = p(xx)
    torch.cuda.synchronize()
    tail0 = n - 4096 - 5   # not a multiple of the vector width: the separate call starts unaligned -> generic kernel
    vt = p(xx[tail0:])
    assert torch.equal(v[tail0:], vt), "tail differs from a separate call"
    o = orc.OracleChebPoly(c, lb, ub)
    for sl in (slice(0, 1000), slice(n - 1000, n)):
        assert np.array_equal(v[sl].cpu().numpy(), o(xx[sl].cpu().numpy())), "oracle mismatch"
    bad_at = n - 3
    if N > 1:
        x[bad_at, 1] = 1.5
    else:
        x[bad_at, 0] = 1.5
    try:
        p(xx)
        raise SystemExit("out-of-domain point not reported")
    except fastcheb.ArgumentError as e:
        assert e.index == bad_at, (e.index, bad_at)
    print(f"{dims}: {n} points in one call ok; tail bitwise equal; oracle parity at both ends; offender index {bad_at} reported")
    del x, v, vt, xx
    torch.cuda.empty_cache()

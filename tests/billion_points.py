#!/usr/bin/env python
"""One-off check of BASELINE configs[2]'s full size on one GPU: 2^30 (1.07e9) device-resident points of the 3-D
order-(32,32,32) SVector{3} polynomial in ONE call (25.8 GB in, 25.8 GB out; 64-bit indexing), compared with the same
points evaluated in 64 separate calls (bitwise) and with the CPU oracle at both ends of the batch."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cheb_oracle as orc  # noqa: E402
import fastcheb  # noqa: E402

rng = np.random.default_rng(4)
coefs = rng.uniform(-1, 1, (33, 33, 33, 3))
lb, ub = -np.ones(3), np.ones(3)
c = fastcheb.ChebPoly(coefs, lb, ub)
n = 1 << 30
g = torch.Generator(device="cuda")
g.manual_seed(5)
pts = torch.empty((n, 3), dtype=torch.float64, device="cuda")
for i in range(16):                                   # generate in pieces (torch.rand temporaries)
    s = slice(i * (n // 16), (i + 1) * (n // 16))
    pts[s] = torch.rand((n // 16, 3), dtype=torch.float64, device="cuda", generator=g) * 2 - 1
torch.cuda.synchronize()
t0 = time.perf_counter()
v = c(pts)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{n} points in one call: {dt:.3f} s = {n / dt:.4e} points/s")
ok = True
for i in range(64):
    s = slice(i * (n // 64), (i + 1) * (n // 64))
    ok = ok and torch.equal(c(pts[s]), v[s])
print("64 separate calls bitwise equal:", ok)
tol = 16 * np.finfo(float).eps * np.abs(coefs).sum()
ref = orc.OracleChebPoly(coefs, lb, ub)
for s in (slice(0, 2000), slice(n - 2000, n), slice(n // 2 + 12345, n // 2 + 14345)):
    d = np.abs(v[s].cpu().numpy() - ref(pts[s].cpu().numpy())).max()
    print(f"max |gpu - oracle| on points {s.start}..{s.stop}: {d:.3e} (tol {tol:.3e})")
    ok = ok and d <= tol
print("OK" if ok else "FAILED")

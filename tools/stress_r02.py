#!/usr/bin/env python
"""One-off randomized stress of the round-2 code paths (not part of the test suite; run on a GPU box):
  (a) 1-D evaluation under AUTO for EVERY n in 40..512 (all row counts / tail lengths / short last rows of the product
      basis), 1-4 components, both precisions, coefficient decay rates from resolved to nearly dense: whatever the plan-time
      self-check lets through must stay within the tolerance of the Clenshaw kernel on 9000 points (1500 of them clustered
      at the end points);
  (b) batches of small N-d arrays taken several per CTA (random shapes, ragged tails), against single-array fits;
  (c) shared-points eval_many on random shapes, against the per-polynomial kernel.
Prints the worst ratios to the tolerances; exits non-zero on a violation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fastcheb as fc

rng = np.random.default_rng(int(os.environ.get("STRESS_SEED", "0")))
worst = {"pb_v": 0.0, "pb_J": 0.0, "sh_v": 0.0, "sh_d": 0.0}
fail = 0
nv = nJ = 0

# (a)
for n in range(40, 513):
    dt = np.float32 if rng.integers(0, 4) == 0 else np.float64
    K = [None, None, 2, 3, 4][int(rng.integers(0, 5))]
    lb, ub = float(rng.uniform(-3, 0)), float(rng.uniform(0.5, 4))
    decay = rng.uniform(0.80, 0.97) ** np.arange(n)
    c = rng.uniform(-1, 1, (n,) + ((K,) if K else ())) * (decay if K is None else decay[:, None])
    c = c.astype(dt)
    p = fc.ChebPoly(c, np.array([lb], dtype=dt), np.array([ub], dtype=dt))
    x = (lb + (ub - lb) * rng.random(9000)).astype(dt)
    m = 1500
    d = (ub - lb) * 10.0 ** rng.uniform(-14, -2, m)
    x[:m] = np.where(rng.random(m) < 0.5, lb + d, ub - d).astype(dt)
    x[m], x[m + 1] = lb, ub
    x = np.clip(x, dt(lb), dt(ub))
    p.set_algo(fc.ALGO_CLENSHAW, dt)
    v0, J0 = fc.chebjacobian(p, x)
    # AUTO: the plan-time self-check decides (a large batch triggers the plan); the comparison below is made only for what
    # AUTO actually routes through the product basis
    p.set_algo(fc.ALGO_AUTO, dt)
    v1 = p(x)
    took_v = p.get_algo(False, dt) == fc.ALGO_PRODUCT
    took_J = p.get_algo(True, dt) == fc.ALGO_PRODUCT
    nv += took_v; nJ += took_J
    v1j, J1 = fc.chebjacobian(p, x)
    if not took_v:
        assert np.array_equal(v1, v0)
    if not took_J:
        J1 = J0
    eps = np.finfo(dt).eps
    cw = np.abs(c.astype(np.float64)).reshape(n, -1)
    tol = 16 * eps * cw.sum(axis=0)                                   # per component
    k2 = (np.arange(n, dtype=np.float64) ** 2)[:, None]
    tolJ = 16 * eps * (cw * k2).sum(axis=0) * 2 / (ub - lb)
    rv = (np.abs(v1.astype(np.float64) - v0.astype(np.float64)).reshape(len(x), -1).max(axis=0) / tol).max()
    rJ = (np.abs(J1.astype(np.float64) - J0.astype(np.float64)).reshape(len(x), -1).max(axis=0) / tolJ).max()
    worst["pb_v"], worst["pb_J"] = max(worst["pb_v"], rv), max(worst["pb_J"], rJ)
    if rv > 1 or rJ > 1:
        print("PB VIOLATION n", n, "K", K, np.dtype(dt).name, rv, rJ); fail += 1
    p.close()
print("product basis under AUTO, n = 40..512: taken for values", nv, "of 473, for derivatives", nJ, "; worst value / derivative ratio to tolerance", worst["pb_v"], worst["pb_J"])

# (b)
for it in range(24):
    N = int(rng.integers(1, 4))
    while True:
        dims = tuple(int(v) for v in rng.integers(2, 34, N))
        K = [None, None, 2, 3][int(rng.integers(0, 4))]
        cplx = bool(rng.integers(0, 4) == 0)
        dt = np.float32 if rng.integers(0, 2) else np.float64
        E = (K or 1) * (2 if cplx else 1)
        if N >= 2 and int(np.prod(dims)) * E * np.dtype(dt).itemsize <= 18 * 1024:
            break
        if N == 1:
            N = 2
    howmany = int(rng.integers(2400, 5200))
    shape = (howmany,) + dims + ((K,) if K else ())
    vals = rng.uniform(-1, 1, shape)
    if cplx:
        vals = vals + 1j * rng.uniform(-1, 1, shape)
    vals = vals.astype((np.complex64 if dt == np.float32 else np.complex128) if cplx else dt)
    got, mx = fc.chebcoefs(vals, ndim=len(dims), howmany=howmany, return_maxabs=True)
    for i in sorted({0, 1, howmany // 3, howmany - 2, howmany - 1}):
        one = fc.chebcoefs(vals[i], ndim=len(dims))
        if not np.array_equal(one, got[i]):
            print("BLOCK MISMATCH", dims, K, cplx, np.dtype(dt).name, howmany, i, np.abs(one - got[i]).max()); fail += 1
    a = np.abs(got.astype(np.complex128 if cplx else np.float64)).reshape(howmany, -1, (K or 1)).max(axis=2).max(axis=1) if K else \
        np.abs(got.astype(np.complex128 if cplx else np.float64)).reshape(howmany, -1).max(axis=1)
    if not np.allclose(mx, a, rtol=1e-6 if dt == np.float32 else 1e-13, atol=0):
        print("BLOCK MAXABS MISMATCH", dims, K, cplx, np.dtype(dt).name, howmany); fail += 1
print("small-array blocks: 24 random batches checked against single-array fits")

# (c)
for it in range(int(os.environ.get("STRESS_SHARED", "40"))):
    n = int(rng.integers(1, 66))
    howmany, M = int(rng.integers(1, 900)), int(rng.integers(1, 1600))
    dt = np.float32 if rng.integers(0, 2) else np.float64
    K = [None, None, 2, 3][int(rng.integers(0, 4))]
    cplx = bool(rng.integers(0, 4) == 0)
    shape = (howmany, n) + ((K,) if K else ())
    decay = (0.85 ** np.arange(n)).reshape((1, n) + ((1,) if K else ()))
    c = rng.uniform(-1, 1, shape) * decay
    if cplx:
        c = c + 1j * rng.uniform(-1, 1, shape) * decay
    c = c.astype((np.complex64 if dt == np.float32 else np.complex128) if cplx else dt)
    lb, ub = dt(-1.5), dt(0.75)
    x = np.clip((lb + (ub - lb) * rng.random(M)).astype(dt), lb, ub)
    x[0] = lb
    x[-1] = ub
    if os.environ.get("STRESS_VERBOSE"):
        print("shared cfg", n, howmany, M, K, cplx, np.dtype(dt).name, flush=True)
    v, d = fc.eval_many(c, lb, ub, x, deriv=True)                                  # shared points: the GEMM
    v0, d0 = fc.eval_many(c, lb, ub, np.broadcast_to(x, (howmany, M)).copy(), deriv=True)  # per-polynomial Clenshaw (bit-exact)
    eps = np.finfo(dt).eps
    # strict reading: per polynomial AND per real component (real and imaginary parts separately)
    cr = c.view(dt).reshape(howmany, n, -1).astype(np.float64) if cplx else c.reshape(howmany, n, -1).astype(np.float64)
    ca = np.abs(cr)
    tol = 16 * eps * ca.sum(axis=1)                                                # (howmany, Ereal)
    k2 = (np.arange(n, dtype=np.float64) ** 2)[None, :, None]
    told = 16 * eps * (ca * k2).sum(axis=1) * 2 / (float(ub) - float(lb))
    real = lambda a: (a.view(dt) if cplx else a).reshape(howmany, M, -1).astype(np.float64)
    rv = (np.abs(real(v) - real(v0)).max(axis=1) / np.maximum(tol, 1e-300)).max()
    rd = (np.abs(real(d) - real(d0)).max(axis=1) / np.maximum(told, 1e-300)).max() if n > 1 else 0.0
    worst["sh_v"], worst["sh_d"] = max(worst["sh_v"], rv), max(worst["sh_d"], rd)
    if rv > 1 or rd > 1:
        print("SHARED VIOLATION", n, howmany, M, K, cplx, np.dtype(dt).name, rv, rd); fail += 1
    elif rd > 0.6 or rv > 0.6:
        print("shared: close", n, howmany, M, K, cplx, np.dtype(dt).name, rv, rd)
print("shared points: worst value / derivative ratio to tolerance", worst["sh_v"], worst["sh_d"])
# (d) random N-d / long-line fit batches through the resident kernel and its slab / tile path, against scipy's DCT-I
#     (pocketfft REDFT00 + the normalisation of interp.jl:49-54) and against single-array fits
import scipy.fft

def ref_chebcoefs(v, ndim):
    v = np.asarray(v, dtype=np.complex128 if np.iscomplexobj(v) else np.float64)
    axes = [d for d in range(ndim) if v.shape[d] > 1]
    def tr(a):
        return scipy.fft.dctn(a, type=1, axes=axes) if axes else a.copy()
    c = tr(v.real) + 1j * tr(v.imag) if np.iscomplexobj(v) else tr(v)
    for d in axes:
        nd = v.shape[d]
        c = c / (2 * (nd - 1))
        idx = [slice(None)] * c.ndim
        idx[d] = slice(1, nd - 1)
        c[tuple(idx)] *= 2
    return c

wfit = 0.0
for it in range(int(os.environ.get("STRESS_FITS", "40"))):
    N = int(rng.integers(1, 4))
    dt = np.float32 if rng.integers(0, 2) else np.float64
    K = [None, None, 2, 3][int(rng.integers(0, 4))]
    cplx = bool(rng.integers(0, 4) == 0)
    E = (K or 1) * (2 if cplx else 1)
    while True:
        if N == 1:
            dims = (int(rng.choice([rng.integers(2, 66), rng.integers(66, 130), 129, 201, 257, 97, 81])),)
        else:
            dims = tuple(int(v) for v in rng.choice([1, 2, 3, 5, 9, 17, 20, 33, 40, 64, 65, 129], N))
        nb = int(np.prod(dims)) * E * np.dtype(dt).itemsize
        if nb <= 600 * 1024:
            break
    howmany = int(rng.integers(1, 200 if nb < 40 * 1024 else 12))
    shape = (howmany,) + dims + ((K,) if K else ())
    vals = rng.uniform(-1, 1, shape)
    if cplx:
        vals = vals + 1j * rng.uniform(-1, 1, shape)
    vals = vals.astype((np.complex64 if dt == np.float32 else np.complex128) if cplx else dt)
    if os.environ.get("STRESS_VERBOSE"):
        print("fit cfg", dims, K, cplx, np.dtype(dt).name, howmany, flush=True)
    got = fc.chebcoefs(vals, ndim=len(dims), howmany=howmany)
    tol = 1e-13 if dt == np.float64 else 450 * np.finfo(np.float32).eps
    for i in sorted({0, howmany // 2, howmany - 1}):
        want = ref_chebcoefs(vals[i], len(dims))
        r = np.abs(got[i] - want).max() / (tol * np.abs(want).max())
        wfit = max(wfit, r)
        if r > 1:
            print("FIT VIOLATION", dims, K, cplx, np.dtype(dt).name, howmany, i, r); fail += 1
        one = fc.chebcoefs(vals[i], ndim=len(dims))
        if not np.array_equal(one, got[i]):
            d1 = np.abs(one - got[i]).max() / (tol * np.abs(want).max())
            if d1 > 1:
                print("FIT single-vs-batch VIOLATION", dims, K, cplx, np.dtype(dt).name, howmany, i, d1); fail += 1
print("random fit batches: worst ratio to tolerance", wfit)
# (e) N-d evaluation: the tensor-core paths against the Clenshaw kernel with the STRICT per-component tolerance
#     (16 eps sum_k |c_{k,e}| for component e; the test suite sums over all components)
wnd_v = wnd_J = 0.0
for it in range(int(os.environ.get("STRESS_ND", "40"))):
    N = int(rng.integers(2, 5))
    while True:
        dims = tuple(int(v) for v in rng.integers(5, 41, N))
        if np.prod(dims) <= 60000 and np.prod(dims[1:]) >= 8:
            break
    K = [None, 2, 3, 5][int(rng.integers(0, 4))]
    cplx = bool(rng.integers(0, 4) == 0)
    dt = np.float32 if rng.integers(0, 3) == 0 else np.float64
    shape = dims + ((K,) if K else ())
    c = rng.uniform(-1, 1, shape)
    if cplx:
        c = c + 1j * rng.uniform(-1, 1, shape)
    c = c.astype((np.complex64 if dt == np.float32 else np.complex128) if cplx else dt)
    lb = rng.uniform(-2, 0, N).astype(dt)
    ub = (lb + rng.uniform(0.5, 3, N)).astype(dt)
    npts = int(rng.integers(1, 5000))
    pts = np.clip((lb + (ub - lb) * rng.random((npts, N))).astype(dt), lb, ub)
    pts[0], pts[-1] = lb, ub
    p = fc.ChebPoly(c, lb, ub)
    p.set_algo(fc.ALGO_CLENSHAW, dt)
    v0, J0 = fc.chebjacobian(p, pts)
    p.set_algo(fc.ALGO_DMMA, dt)
    if p.get_algo(False, dt) not in (fc.ALGO_DMMA, fc.ALGO_DMMA_RING):
        p.close(); continue
    v1, J1 = fc.chebjacobian(p, pts)
    eps = np.finfo(dt).eps
    cr = (c.view(dt) if cplx else c).reshape(int(np.prod(dims)), -1).astype(np.float64)   # (coefficients, real components)
    tol = 16 * eps * np.abs(cr).sum(axis=0)
    real = lambda a: (np.ascontiguousarray(a).view(dt) if cplx else a)
    rv = (np.abs(real(v1).reshape(npts, -1).astype(np.float64) - real(v0).reshape(npts, -1).astype(np.float64)).max(axis=0) / tol).max()
    wnd_v = max(wnd_v, rv)
    for d in range(N):
        kshape = [1] * N + [1]
        kshape[d] = dims[d]
        k2 = (np.arange(dims[d], dtype=np.float64) ** 2).reshape(kshape)
        ca = np.abs((c.view(dt) if cplx else c).reshape(dims + (-1,)).astype(np.float64))
        tolJ = 16 * eps * (ca * k2).reshape(int(np.prod(dims)), -1).sum(axis=0) * 2 / (float(ub[d]) - float(lb[d]))
        Jd0 = real(np.ascontiguousarray(J0[..., d] if J0.ndim == 3 else J0)).reshape(npts, -1).astype(np.float64)
        Jd1 = real(np.ascontiguousarray(J1[..., d] if J1.ndim == 3 else J1)).reshape(npts, -1).astype(np.float64)
        rJ = (np.abs(Jd1 - Jd0).max(axis=0) / tolJ).max()
        wnd_J = max(wnd_J, rJ)
    if rv > 1 or wnd_J > 1:
        print("ND VIOLATION", dims, K, cplx, np.dtype(dt).name, rv, wnd_J); fail += 1
    p.close()
print("N-d tensor-core paths vs Clenshaw, per-component tolerances: worst value / Jacobian ratio", wnd_v, wnd_J)
print("FAILURES", fail)
sys.exit(1 if fail else 0)

"""Seeded random shapes through every evaluation path vs the oracle: dims 1..40 per dimension, 1-4 dimensions, scalar /
vector / complex, both precisions, odd batch sizes, non-trivial boxes.  Clenshaw must be bit-identical; whatever AUTO
picks (Clenshaw, ring DMMA, resident 2-D DMMA) must be within 16 eps sum|c| (values) and the differentiated-series bound
(Jacobians), as in test_gpu_eval.py."""
import os

import numpy as np
import pytest

from parity_util import jac_tol

pytestmark = pytest.mark.gpu

# other random streams on demand: FASTCHEB_FUZZ_OFFSET=n shifts every seed (the default run is deterministic)
OFFSET = 100000 * int(os.environ.get("FASTCHEB_FUZZ_OFFSET", "0"))


def random_case(rng):
    N = int(rng.integers(1, 5))
    while True:
        dims = tuple(int(v) for v in rng.integers(1, 41, N))
        if np.prod(dims) <= 60000:
            break
    K = [None, None, 2, 3, 5][int(rng.integers(0, 5))]
    cplx = bool(rng.integers(0, 4) == 0)
    dt = np.float32 if rng.integers(0, 3) == 0 else np.float64
    return dims, K, cplx, dt


@pytest.mark.parametrize("seed", range(40))
def test_random_shapes(fc, orc, seed):
    rng = np.random.default_rng(1000 + seed + OFFSET)
    for _ in range(10):
        dims, K, cplx, dt = random_case(rng)
        N = len(dims)
        shape = dims + ((K,) if K else ())
        c = rng.uniform(-1, 1, shape)
        if cplx:
            c = (c + 1j * rng.uniform(-1, 1, shape)).astype(np.complex64 if dt == np.float32 else np.complex128)
        else:
            c = c.astype(dt)
        lb = rng.uniform(-3, 1, N).astype(dt)
        ub = (lb + rng.uniform(0.25, 4, N)).astype(dt)
        npts = int(rng.integers(1, 3000))
        pts = (lb + (ub - lb) * rng.random((npts, N))).astype(dt)
        pts = np.clip(pts, lb, ub)
        pts[0] = lb
        pts[-1] = ub
        ref = orc.OracleChebPoly(c, lb, ub)
        v0, J0 = ref.jacobian(pts)
        tol = 16 * np.finfo(dt).eps * np.abs(c).sum()
        poly = fc.ChebPoly(c, lb, ub)
        tag = (dims, K, cplx, np.dtype(dt).name, npts)
        poly.set_algo(fc.ALGO_CLENSHAW, dt)
        assert np.array_equal(poly(pts), v0), tag
        v, J = fc.chebjacobian(poly, pts)
        assert np.array_equal(v, v0) and np.array_equal(J, J0), tag
        poly.set_algo(fc.ALGO_AUTO, dt)
        assert np.abs(poly(pts) - v0).max() <= tol, tag
        v, J = fc.chebjacobian(poly, pts)
        assert np.abs(v - v0).max() <= tol, tag
        for d, n in enumerate(dims):
            told = jac_tol(c, len(dims), d, lb, ub, dt)
            assert np.abs(J[:, :, d] - J0[:, :, d]).max() <= told, (tag, d)
        if N >= 2 and dims[0] >= 5 and np.prod(dims[1:]) >= 8 and N <= 4:
            poly.set_algo(fc.ALGO_DMMA, dt)       # forced tensor-core path on shapes AUTO would leave to Clenshaw
            assert np.abs(poly(pts) - v0).max() <= tol, tag
            v, J = fc.chebjacobian(poly, pts)
            assert np.abs(v - v0).max() <= tol, tag
            for d, n in enumerate(dims):
                told = jac_tol(c, len(dims), d, lb, ub, dt)
                assert np.abs(J[:, :, d] - J0[:, :, d]).max() <= told, (tag, d)
        poly.close()


@pytest.mark.parametrize("seed", range(10))
def test_random_fits(fc, orc, seed):
    """chebcoefs on random shapes: batches of 1-D lines of every length 2..65 (tensor-core path incl. ragged tails and
    vector / complex elements) and small N-d arrays (generic path), both precisions, vs the long-double oracle."""
    rng = np.random.default_rng(2000 + seed + OFFSET)
    for _ in range(8):
        dt = np.float32 if rng.integers(0, 2) else np.float64
        tol = 1e-13 if dt == np.float64 else 450 * np.finfo(np.float32).eps
        cplx = bool(rng.integers(0, 3) == 0)
        K = [None, None, 2, 3][int(rng.integers(0, 4))]
        if rng.integers(0, 2):
            n, howmany = int(rng.integers(2, 66)), int(rng.integers(1, 700))
            shape = (howmany, n) + ((K,) if K else ())
            vals = rng.uniform(-1, 1, shape)
            if cplx:
                vals = vals + 1j * rng.uniform(-1, 1, shape)
            vals = vals.astype((np.complex64 if dt == np.float32 else np.complex128) if cplx else dt)
            got = fc.chebcoefs(vals, ndim=1, howmany=howmany)
            wide = vals.astype(np.complex128 if cplx else np.float64)
            want = np.stack([orc.chebcoefs(x, 1) for x in wide])
        else:
            N = int(rng.integers(1, 4))
            dims = tuple(int(v) for v in rng.integers(1, 30, N))
            shape = dims + ((K,) if K else ())
            vals = rng.uniform(-1, 1, shape)
            if cplx:
                vals = vals + 1j * rng.uniform(-1, 1, shape)
            vals = vals.astype((np.complex64 if dt == np.float32 else np.complex128) if cplx else dt)
            got = fc.chebcoefs(vals, ndim=N)
            want = orc.chebcoefs(vals.astype(np.complex128 if cplx else np.float64), N)
        assert got.shape == want.shape and got.dtype == vals.dtype
        assert np.abs(got - want).max() <= tol * max(np.abs(want).max(), 1e-300), (shape, cplx, np.dtype(dt).name)


@pytest.mark.parametrize("seed", range(8))
def test_random_domain_errors(fc, seed):
    """The FIRST offending point (eval.jl:64: outside the closed box, or NaN) is what every kernel family reports,
    wherever it sits in the batch (first / last unit, several bad points, large batches that use all warps)."""
    rng = np.random.default_rng(3000 + seed + OFFSET)
    for _ in range(6):
        dims, K, cplx, dt = random_case(rng)
        N = len(dims)
        shape = dims + ((K,) if K else ())
        c = rng.uniform(-1, 1, shape).astype(dt)
        lb = rng.uniform(-3, 1, N).astype(dt)
        ub = (lb + rng.uniform(0.25, 4, N)).astype(dt)
        npts = int(rng.integers(1, 200_000 if rng.integers(0, 3) == 0 else 3000))
        pts = (lb + (ub - lb) * rng.random((npts, N))).astype(dt)
        pts = np.clip(pts, lb, ub)
        bad_idx = sorted(set(int(v) for v in rng.integers(0, npts, int(rng.integers(1, 4)))))
        for i in bad_idx:
            d = int(rng.integers(0, N))
            pts[i, d] = [np.nan, ub[d] + (ub[d] - lb[d]) * dt(0.01) + dt(1e-3), lb[d] - dt(1e-3)][int(rng.integers(0, 3))]
        poly = fc.ChebPoly(c, lb, ub)
        algos = [fc.ALGO_CLENSHAW, fc.ALGO_AUTO]
        if N >= 2 and N <= 4 and dims[0] >= 5 and np.prod(dims[1:]) >= 8:
            algos += [fc.ALGO_DMMA, fc.ALGO_DMMA_RING]
        for algo in algos:
            poly.set_algo(algo, dt)
            for jac in (False, True):
                with pytest.raises(fc.ArgumentError):
                    fc.chebjacobian(poly, pts) if jac else poly(pts)
                assert fc._capi.error_index() == bad_idx[0], (dims, K, np.dtype(dt).name, npts, algo, jac, bad_idx)
        poly.close()

"""Multi-GPU evaluation and fits from ONE host process through the C ABI (fastcheb_multi_*, include/fastcheb.h):
the points of `c.(ys)` are block-partitioned over every GPU of the box, results come back in one array in point
order, and the ArgumentError names the first offending point of the WHOLE batch (eval.jl:64).  Runs with however
many GPUs the box has (1 on the single-GPU tier: the same code path with one worker)."""
import ctypes

import numpy as np
import pytest

from parity_util import jac_tol

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def _ndev(fc):
    return fc._capi.lib().fastcheb_device_count()


@pytest.mark.parametrize("case", [((33, 33, 33), 3, False), ((65, 65), None, False), ((201,), None, False), ((9, 8, 7, 6), None, True)],
                         ids=["3d-K3", "2d", "1d", "4d-complex"])
def test_multi_eval_host_matches_oracle_and_single_gpu(fc, orc, case):
    dims, K, cplx = case
    rng = np.random.default_rng(31)
    N = len(dims)
    shape = dims + ((K,) if K else ())
    c = rng.uniform(-1, 1, shape) + (1j * rng.uniform(-1, 1, shape) if cplx else 0)
    lb, ub = -1.0 - 0.25 * np.arange(N), 0.5 + 0.75 * np.arange(N)
    npts = 40_003                                        # not a multiple of the device count: ragged blocks
    pts = lb + (ub - lb) * rng.random((npts, N))
    pts[0], pts[-1] = lb, ub
    x = pts if N > 1 else pts[:, 0]
    m = fc.MultiChebPoly(c, lb, ub)
    assert len(m.devices) == _ndev(fc)
    single = fc.ChebPoly(c, lb, ub)
    v = m(x)
    assert np.array_equal(v, single(x))                   # a point's result does not depend on which GPU computed it
    v0, J0 = orc.OracleChebPoly(c, lb, ub).jacobian(x)
    assert np.abs(v - v0).max() <= 16 * EPS * np.abs(c).sum()
    v2, J = fc.chebjacobian(m, x)
    assert np.array_equal(v2, fc.chebjacobian(single, x)[0])
    for d in range(N):
        assert np.abs(J[:, :, d] - J0[:, :, d]).max() <= max(jac_tol(c, N, d, lb, ub), 0.0)
    # the first offending point of the whole batch, wherever its block lives
    bad = pts.copy()
    bad[npts - 5, N - 1] = ub[N - 1] + 1.0                # in the last device's block
    bad[npts // 2 + 1, 0] = np.nan                        # in a middle block: this one is first
    with pytest.raises(fc.ArgumentError) as ei:
        m(bad if N > 1 else bad[:, 0])
    assert ei.value.index == npts // 2 + 1
    with pytest.raises(fc.ArgumentError) as ei:
        fc.chebjacobian(m, bad if N > 1 else bad[:, 0])
    assert ei.value.index == npts // 2 + 1
    m.close()


def test_multi_eval_device_resident(fc, orc):
    """Points and results on devices[0]; the other GPUs read / write them through peer access (NVLink)."""
    import torch

    rng = np.random.default_rng(32)
    c = rng.uniform(-1, 1, (17, 16, 15, 2))
    lb, ub = -np.ones(3), np.ones(3)
    pts = rng.uniform(-1, 1, (200_001, 3))
    m = fc.MultiChebPoly(c, lb, ub)
    t = torch.from_numpy(pts).cuda(m.devices[0])
    v = m(t)
    ref = orc.OracleChebPoly(c, lb, ub)(pts)
    assert v.is_cuda and np.abs(v.cpu().numpy() - ref).max() <= 16 * EPS * np.abs(c).sum()
    v2, J = fc.chebjacobian(m, t)
    assert torch.equal(v2, v) and J.shape == (len(pts), 2, 3)
    J0 = orc.OracleChebPoly(c, lb, ub).jacobian(pts)[1]
    for d in range(3):
        assert np.abs(J.cpu().numpy()[:, :, d] - J0[:, :, d]).max() <= jac_tol(c, 3, d, lb, ub)
    bad = t.clone()
    bad[150_000, 1] = 2.0
    with pytest.raises(fc.ArgumentError) as ei:
        m(bad)
    assert ei.value.index == 150_000
    m.close()


def test_multi_tuple_layout_and_gradient_error(fc, orc):
    rng = np.random.default_rng(33)
    c = rng.uniform(-1, 1, (7, 6, 2))
    lb, ub = np.zeros(2), np.ones(2)
    pts = rng.random((1001, 2))
    m = fc.MultiChebPoly(c, lb, ub)
    h = m._handle(np.float64)
    L = fc._capi.lib()
    for i in range(L.fastcheb_multi_ngpus(h)):
        assert L.fastcheb_set_algo(L.fastcheb_multi_poly(h, i), fc.ALGO_CLENSHAW) == 0
    rec = np.empty((1001, 2 + 4))
    assert L.fastcheb_multi_eval_jac_tuple(h, pts.ctypes.data, 1001, rec.ctypes.data, fc._capi.MEM_HOST) == 0
    v0, J0 = orc.OracleChebPoly(c, lb, ub).jacobian(pts)
    assert np.array_equal(rec[:, :2], v0)
    assert np.array_equal(rec[:, 2:].reshape(1001, 2, 2), np.swapaxes(J0, 1, 2))   # Julia's Tuple{SVector, SMatrix} bytes
    with pytest.raises(fc.DimensionMismatch):
        fc.chebgradient(m, pts)                            # eval.jl:170
    m.close()


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_multi_fit(fc, orc, dt):
    rng = np.random.default_rng(34)
    vals = rng.uniform(-1, 1, (10_007, 65)).astype(dt)
    got, mx = fc.chebcoefs_multi(vals, return_maxabs=True)
    one = fc.chebcoefs(vals, ndim=1, howmany=len(vals))
    tol = 1e-13 if dt == np.float64 else 450 * np.finfo(np.float32).eps
    # the ragged tail of every device's block takes another kernel than the body (same transform, other rounding), so
    # the split changes last bits at the block boundaries; with one device the partition is the single-GPU one
    assert np.abs(got - one).max() <= tol * np.abs(one).max()
    if _ndev(fc) == 1:
        assert np.array_equal(got, one)
    want = np.stack([orc.chebcoefs(v.astype(np.float64), 1) for v in vals[:50]])
    assert np.abs(got[:50] - want).max() <= tol * np.abs(want).max()
    assert np.allclose(mx[:50], np.abs(want).max(axis=1), rtol=1e-5 if dt == np.float32 else 1e-12)
    bad = vals.copy()
    bad[9000, 7] = np.inf
    with pytest.raises(fc.DomainError):
        fc.chebcoefs_multi(bad)
    assert fc._capi.error_index() == 9000 * 65 + 7


def test_multi_bad_arguments(fc):
    L = fc._capi.lib()
    h = ctypes.c_void_p()
    c = np.ones(4)
    args = (1, fc._capi.i64_array((4,)), fc._capi.F64, 0, 1, c.ctypes.data, fc._capi.MEM_HOST, fc._capi.dbl_array([0.0]), fc._capi.dbl_array([1.0]))
    assert L.fastcheb_multi_create(ctypes.byref(h), _ndev(fc) + 1, None, *args) == fc._capi.EINVAL
    dup = (ctypes.c_int * 2)(0, 0)
    if _ndev(fc) >= 2:
        assert L.fastcheb_multi_create(ctypes.byref(h), 2, dup, *args) == fc._capi.EINVAL
    assert L.fastcheb_multi_destroy(None) == 0

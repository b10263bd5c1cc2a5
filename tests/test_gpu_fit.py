"""GPU parity: the DCT-I fit (chebcoefs), droptol and chebinterp through the C ABI vs the oracle.
Tolerance (north_star): |dc| <= 1e-13 max|c| per fitted Float64 coefficient; the same multiple of eps
(450 eps) for Float32."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [((201,), None), ((49,), None), ((65,), 3), ((2,), None), ((1,), None), ((65, 65), None), ((49, 40), None),
          ((11, 21), 2), ((33, 33, 33), None), ((9, 1, 7), None), ((2, 3, 2, 5), 2), ((10001,), None)]


def tol_for(dt):
    return 1e-13 if dt == np.float64 else 450 * np.finfo(np.float32).eps


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("shape,K", SHAPES, ids=lambda v: str(v))
def test_chebcoefs_parity(fc, orc, shape, K, dt):
    rng = np.random.default_rng(21)
    full = shape + ((K,) if K else ())
    vals = rng.uniform(-1, 1, full).astype(dt)
    want = orc.chebcoefs(vals.astype(np.float64), len(shape))     # long-double truth, rounded to double
    got = fc.chebcoefs(vals, ndim=len(shape))
    assert got.dtype == dt and got.shape == want.shape
    assert np.abs(got - want).max() <= tol_for(dt) * np.abs(want).max()


def test_chebcoefs_complex_and_batched(fc, orc):
    rng = np.random.default_rng(22)
    v = rng.uniform(-1, 1, (17, 9)) + 1j * rng.uniform(-1, 1, (17, 9))
    got = fc.chebcoefs(v, ndim=2)
    want = orc.chebcoefs(v, 2)
    assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max()
    for dt in (np.float64, np.float32):
        vals = rng.uniform(-1, 1, (1000, 65)).astype(dt)
        got = fc.chebcoefs(vals, ndim=1, howmany=1000)
        want = np.stack([orc.chebcoefs(x.astype(np.float64), 1) for x in vals])
        assert np.abs(got - want).max() <= tol_for(dt) * np.abs(want).max()
    for n in (5, 17, 33, 34, 64, 129, 257):
        vals = rng.uniform(-1, 1, (37, n))
        got = fc.chebcoefs(vals, ndim=1, howmany=37)
        want = np.stack([orc.chebcoefs(x, 1) for x in vals])
        assert np.abs(got - want).max() <= 1e-13 * np.abs(want).max(), n


def test_nonfinite_raises_domainerror(fc):
    v = np.ones((9, 9))
    v[3, 4] = np.inf
    with pytest.raises(fc.DomainError):
        fc.chebcoefs(v)                           # interp.jl:40
    v[3, 4] = np.nan
    with pytest.raises(fc.DomainError):
        fc.chebinterp(v, [0, 0], [1, 1])


def test_droptol_and_chebinterp(fc, orc):
    rng = np.random.default_rng(23)
    lb, ub = np.array([1.0, 3.0]), np.array([2.0, 4.0])
    x = fc.chebpoints((10, 20), lb, ub)
    g = np.sin(x[..., 0] + np.cos(x[..., 1]))
    for tol in (None, 0.0, 0.01, 1e-8):
        a = fc.chebinterp(g, lb, ub, tol=tol)
        b = orc.chebinterp(g, lb, ub, tol=tol)
        assert a.coefs.shape == b.coefs.shape, tol
        assert np.abs(a.coefs - b.coefs).max() <= 1e-13 * np.abs(b.coefs).max()
    c = rng.uniform(-1, 1, (8, 7, 6, 2)) * (10.0 ** -rng.integers(0, 18, (8, 7, 6, 1)))
    for tol in (0.0, 1e-15, 1e-10, 1e-3):
        assert fc.droptol(c, tol, ndim=3).shape == orc.droptol(c, 3, tol).shape
    cc = (rng.uniform(-1, 1, (9, 5)) + 1j * rng.uniform(-1, 1, (9, 5))) * 10.0 ** -np.arange(9)[:, None]
    for tol in (1e-12, 1e-6, 1e-3):
        assert fc.droptol(cc, tol).shape == orc.droptol(cc, 2, tol).shape   # complex modulus (interp.jl:74)


def test_fit_then_eval_roundtrip(fc):
    """Size-independent property: the interpolant reproduces its data on the Chebyshev grid."""
    lb, ub = np.array([-1.0, 0.0, 2.0]), np.array([1.0, 1.0, 5.0])
    x = fc.chebpoints((20, 15, 12), lb, ub)
    f = np.stack([np.sin(x[..., 0] * x[..., 1]) + x[..., 2], np.cos(x[..., 2]) * x[..., 0]], -1)
    c = fc.chebinterp(f, lb, ub, tol=0.0)
    got = c(np.clip(x.reshape(-1, 3), lb, ub))
    assert np.abs(got - f.reshape(-1, 2)).max() <= 1e-12

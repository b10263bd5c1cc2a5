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


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("kind", ["K3", "complex", "K2complex"])
def test_batched_order64_vector_and_complex(fc, orc, dt, kind):
    """Batched order-64 fits with E = 3, 2 and 4 inline reals per element: units with dead lines (E = 3 -> 2 fits =
    6 of 8 lines) and strided lines on the tensor-core path, incl. the Float32 slot maps; non-finite detection."""
    rng = np.random.default_rng(29)
    howmany = 501
    if kind == "K3":
        vals = rng.uniform(-1, 1, (howmany, 65, 3)).astype(dt)
    else:
        shape = (howmany, 65) + ((2,) if kind == "K2complex" else ())
        vals = (rng.uniform(-1, 1, shape) + 1j * rng.uniform(-1, 1, shape)).astype(np.complex128 if dt == np.float64 else np.complex64)
    got = fc.chebcoefs(vals, ndim=1, howmany=howmany)
    assert got.shape == vals.shape and got.dtype == vals.dtype
    wide = vals.astype(np.complex128 if np.iscomplexobj(vals) else np.float64)
    want = np.stack([orc.chebcoefs(x, 1) for x in wide])
    assert np.abs(got - want).max() <= tol_for(dt) * np.abs(want).max()
    bad = vals.copy()
    bad[417, 33] = np.nan
    with pytest.raises(fc.DomainError):
        fc.chebcoefs(bad, ndim=1, howmany=howmany)


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


@pytest.mark.parametrize("seed", range(6))
def test_chebinterp_fused_droptol_matches_separate_calls(fc, seed):
    """chebinterp's trimming statistics come from the resident fit kernel itself (fit_nd.cuh: max / min of the element
    norms, then every element >= tol * max marks its hyperplanes); chebcoefs followed by droptol goes through the
    separate reduction kernels.  Same coefficients, same decisions: shapes and coefficients must agree exactly, for random
    shapes (size-1 dimensions, vector / complex elements), decay rates and tolerances."""
    rng = np.random.default_rng(4000 + seed)
    for _ in range(12):
        N = int(rng.integers(1, 4))
        dims = tuple(int(v) for v in rng.choice([1, 2, 3, 5, 9, 17, 20, 33, 40, 65], N))
        if int(np.prod(dims)) > 60000:
            dims = dims[:-1] + (5,)
        dt = np.float32 if rng.integers(0, 3) == 0 else np.float64
        cplx = bool(rng.integers(0, 4) == 0)
        K = [None, None, 2, 3][int(rng.integers(0, 4))]
        lb, ub = -np.ones(N), np.ones(N) * 2
        x = fc.chebpoints(tuple(d - 1 for d in dims), lb, ub) if N > 1 else fc.chebpoints(dims[0] - 1, lb[0], ub[0])[..., None]
        freq = rng.uniform(0.2, 6.0, N)
        g = np.ones(dims)
        for d in range(N):
            g = g * np.cos(freq[d] * x[..., d] + rng.uniform(0, 3))
        if K:
            g = np.stack([g * (10.0 ** -int(rng.integers(0, 6))) + 0.1 * k for k in range(K)], axis=-1)
        if cplx:
            g = g + 1j * np.roll(g, 1, axis=0) * 10.0 ** -int(rng.integers(0, 9))
        g = g.astype((np.complex64 if dt == np.float32 else np.complex128) if cplx else dt)
        tol = [None, 0.0, 1e-2, 1e-5, 1e-9, 1e-14][int(rng.integers(0, 6))]
        a = fc.chebinterp(g, lb, ub, tol=tol)
        c = fc.chebcoefs(g, ndim=N)
        eff = (np.finfo(dt).eps if tol is None else tol)
        want = fc.droptol(c, eff, ndim=N)
        assert a.coefs.shape == want.shape, (dims, K, cplx, np.dtype(dt).name, tol)
        assert np.array_equal(a.coefs, want), (dims, K, cplx, np.dtype(dt).name, tol)
        a.close()


def test_fit_then_eval_roundtrip(fc):
    """Size-independent property: the interpolant reproduces its data on the Chebyshev grid."""
    lb, ub = np.array([-1.0, 0.0, 2.0]), np.array([1.0, 1.0, 5.0])
    x = fc.chebpoints((20, 15, 12), lb, ub)
    f = np.stack([np.sin(x[..., 0] * x[..., 1]) + x[..., 2], np.cos(x[..., 2]) * x[..., 0]], -1)
    c = fc.chebinterp(f, lb, ub, tol=0.0)
    got = c(np.clip(x.reshape(-1, 3), lb, ub))
    assert np.abs(got - f.reshape(-1, 2)).max() <= 1e-12


def _oracle_batch(orc, vals, K):
    return np.stack([orc.chebcoefs(np.asarray(v, dtype=np.complex128 if np.iscomplexobj(v) else np.float64), 1) for v in vals])


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7, 8, 9, 13, 17, 21, 33, 34, 35, 40, 49, 61, 64, 65])
def test_batched_tensor_core_fit(fc, orc, n, dt):
    """configs[4]: many independent short 1-D fits (fit_dmma.cuh) incl. a ragged tail that the generic
    kernel finishes; scalar, K-vector and complex elements."""
    rng = np.random.default_rng(100 + n)
    for K, cplx, howmany in ((None, False, 531), (3, False, 203), (2, True, 150), (None, True, 300)):
        shape = (howmany, n) + ((K,) if K else ())
        vals = rng.uniform(-1, 1, shape)
        if cplx:
            vals = vals + 1j * rng.uniform(-1, 1, shape)
            vals = vals.astype(np.complex128 if dt == np.float64 else np.complex64)
        else:
            vals = vals.astype(dt)
        got, mx = fc.chebcoefs(vals, ndim=1, howmany=howmany, return_maxabs=True)
        want = _oracle_batch(orc, vals, K)
        assert got.shape == want.shape and got.dtype == vals.dtype
        assert np.abs(got - want).max() <= tol_for(dt) * np.abs(want).max(), (K, cplx)
        wmx = np.abs(got.reshape(howmany, -1)).max(axis=1)
        assert np.allclose(mx, wmx, rtol=1e-6 if dt == np.float32 else 1e-14, atol=0)


def test_batched_fit_nonfinite_index(fc):
    vals = np.ones((4096, 65))
    vals[1234, 17] = np.nan
    vals[3000, 2] = np.inf
    with pytest.raises(fc.DomainError) as ei:
        fc.chebcoefs(vals, ndim=1, howmany=4096)
    assert str(1234 * 65 + 17) in str(ei.value)          # first offending flat index (interp.jl:40)
    vals = np.ones((4099, 65))                              # ... and in the ragged tail
    vals[4098, 64] = np.nan
    with pytest.raises(fc.DomainError) as ei:
        fc.chebcoefs(vals, ndim=1, howmany=4099)
    assert str(4098 * 65 + 64) in str(ei.value)


def test_batched_fit_linearity_and_constants(fc):
    """Size-independent properties at the full configs[4] size: DCT-I is linear; a constant line has the single
    coefficient c0 = value; T_k sampled on the grid gives the unit vector e_k."""
    rng = np.random.default_rng(7)
    M, n = 100_000, 65
    a, b = rng.uniform(-1, 1, (M, n)), rng.uniform(-1, 1, (M, n))
    ca, cb, cab = (fc.chebcoefs(v, ndim=1, howmany=M) for v in (a, b, 2.0 * a - 0.5 * b))
    assert np.abs(cab - (2.0 * ca - 0.5 * cb)).max() <= 1e-13 * np.abs(cab).max()
    k = rng.integers(0, n, M)
    j = np.arange(n)
    grid = np.cos(np.pi * j / (n - 1))                      # interp.jl:6: index 0 is the upper bound
    tk = np.cos(k[:, None] * np.arccos(grid)[None, :])
    ck = fc.chebcoefs(tk, ndim=1, howmany=M)
    want = np.zeros((M, n))
    want[np.arange(M), k] = 1.0
    assert np.abs(ck - want).max() <= 1e-13


def test_large_host_batch_runs_in_chunks_on_pooled_streams(fc, orc):
    """Host batches above 64 MB are cut into chunks on two internal streams (fit.cu; the streams come from the per-device
    pool, so a second call reuses them): every fit equals the same fit made alone, and the per-fit maxima are right."""
    rng = np.random.default_rng(99)
    howmany, n = 150_000, 65                                  # 78 MB of Float64
    vals = rng.uniform(-1, 1, (howmany, n))
    for _ in range(2):
        c, mx = fc.chebcoefs(vals, ndim=1, howmany=howmany, return_maxabs=True)
        assert c.shape == vals.shape
        for i in (0, 1, 37_499, 37_500, 74_999, 75_000, 112_500, 149_999):
            ref = orc.chebcoefs(vals[i], 1)
            assert np.abs(c[i] - ref).max() <= 1e-13 * np.abs(ref).max()
            assert np.isclose(mx[i], np.abs(c[i]).max(), rtol=1e-14, atol=0)
    bad = vals.copy()
    bad[120_000, 7] = np.inf
    with pytest.raises(fc.DomainError):
        fc.chebcoefs(bad, ndim=1, howmany=howmany)

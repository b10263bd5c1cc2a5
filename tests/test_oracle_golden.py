"""CPU: pin the oracle against every known answer the reference publishes for the hot path
(SURVEY.md §8c): README.md:65-146 and test/runtests.jl testsets "1d test", "2d test", "degree 1",
"degree 0", "inference", "boundscheck sensitivity".  Also cross-checks the oracle's three grades of
restatement against each other (C Clenshaw vs long-double truth, long-double DCT vs pocketfft)."""
import numpy as np
import pytest

import known_answers as ka


@pytest.fixture(scope="module")
def api(orc):
    return ka.Api(
        chebpoints=orc.chebpoints,
        interp=lambda vals, lb, ub, tol: orc.chebinterp(vals, lb, ub, tol=tol),
        value=lambda c, x: c(x),
        grad=lambda c, x: c.gradient(x),
        jac=lambda c, x: c.jacobian(x),
        interp_v1=lambda vals, lb, ub: orc.chebinterp_v1(vals, lb, ub),
        coefs=lambda c: c.coefs,
        ArgumentError=ValueError, DomainError=FloatingPointError, DimensionMismatch=TypeError)


@pytest.mark.parametrize("check", ka.ALL_CHECKS, ids=lambda f: f.__name__)
def test_reference_known_answers(api, check):
    check(api)


@pytest.mark.parametrize("T", [np.float32, np.float64], ids=["f32", "f64"])
@pytest.mark.parametrize("check", ka.TYPED_CHECKS, ids=lambda f: f.__name__)
def test_reference_testsets(api, check, T):
    check(api, T)


def test_fft_restatement_matches_long_double_fit(orc):
    """pocketfft DCT-I (stand-in for FFTW REDFT00, interp.jl:44) == the long-double direct cosine sum."""
    rng = np.random.default_rng(1)
    for shape in [(201,), (49, 40), (11, 21), (9, 8, 7), (1, 5), (2, 1, 3), (33, 33, 3)]:
        ndim = len(shape) if shape != (33, 33, 3) else 2
        v = rng.uniform(-1, 1, shape)
        a, b = orc.chebcoefs(v, ndim), orc.chebcoefs_fft(v, ndim)
        assert np.abs(a - b).max() <= 1e-14 * np.abs(a).max()
    vc = rng.uniform(-1, 1, (17, 9)) + 1j * rng.uniform(-1, 1, (17, 9))
    assert np.abs(orc.chebcoefs(vc, 2) - orc.chebcoefs_fft(vc, 2)).max() <= 1e-14


def test_clenshaw_matches_long_double_truth(orc):
    """C restatement of eval.jl:16-56 vs numpy.longdouble truth, within 16 eps sum|c| (north_star tolerance)."""
    rng = np.random.default_rng(2)
    for dims, K, cplx in [((33,), None, False), ((17, 9), None, False), ((9, 8, 7), 3, False), ((5, 4, 3, 6), 2, True),
                          ((1, 7), None, False), ((2, 2), 2, False)]:
        shape = dims + ((K,) if K else ())
        c = rng.uniform(-1, 1, shape)
        if cplx:
            c = c + 1j * rng.uniform(-1, 1, shape)
        N = len(dims)
        lb, ub = -np.arange(1, N + 1, dtype=float), np.arange(2, N + 2, dtype=float)
        pts = lb + (ub - lb) * rng.random((200, N))
        pts[0], pts[1] = lb, ub
        p = orc.OracleChebPoly(c, lb, ub)
        z = (pts - lb) * 2 / (ub - lb) - 1
        truth = orc.eval_truth(c, lb, ub, pts, z=z)
        got = p(pts)
        tol = 16 * np.finfo(float).eps * np.abs(c).sum()
        assert np.abs(got - truth.astype(got.dtype)).max() <= tol


def test_jacobian_matches_finite_difference_of_truth(orc):
    rng = np.random.default_rng(3)
    c = rng.uniform(-1, 1, (9, 8, 7, 2))
    lb, ub = np.array([0.0, -1, 2]), np.array([1.0, 1, 5])
    pts = lb + (ub - lb) * (0.1 + 0.8 * rng.random((20, 3)))
    p = orc.OracleChebPoly(c, lb, ub)
    v, J = p.jacobian(pts)
    h = 1e-6
    for d in range(3):
        e = np.zeros(3)
        e[d] = h
        fd = (orc.eval_truth(c, lb, ub, pts + e) - orc.eval_truth(c, lb, ub, pts - e)) / (2 * h)
        assert np.allclose(J[:, :, d], fd.astype(float), rtol=1e-6, atol=1e-6)


def test_domain_and_shape_errors(orc):
    c = orc.OracleChebPoly(np.ones((3, 3)), np.zeros(2), np.ones(2))
    with pytest.raises(ValueError):
        c(np.array([0.5, 1.5]))                      # eval.jl:64
    with pytest.raises(ValueError):
        c(np.array([np.nan, 0.5]))
    with pytest.raises(FloatingPointError):
        orc.chebinterp(np.array([1.0, np.inf, 2.0]), 0, 1)   # interp.jl:40
    cv = orc.OracleChebPoly(np.ones((3, 3, 2)), np.zeros(2), np.ones(2))
    with pytest.raises(TypeError):
        cv.gradient(np.array([0.5, 0.5]))            # eval.jl:170


def test_droptol_semantics(orc):
    """interp.jl:77-93: tol = 0 is the identity; trailing hyperplanes below tol*max are trimmed per dim."""
    c = np.zeros((6, 5))
    c[:3, :2] = 1.0
    c[3, 0] = 1e-20
    assert orc.droptol(c, 2, 0.0).shape == (6, 5)
    assert orc.droptol(c, 2, 1e-16).shape == (3, 2)
    full = np.ones((4, 4))
    assert orc.droptol(full, 2, 0.5) is full or orc.droptol(full, 2, 0.5).shape == (4, 4)


GOLD_NAMES = ["d1", "d1c", "d2", "d2k", "d3", "d4c", "n1", "n2"]


def _gold():
    import os
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "truth_small.npz"))


@pytest.mark.parametrize("name", GOLD_NAMES)
def test_oracle_against_committed_truth(orc, name):
    """tests/golden/truth_small.npz (made by tests/golden/make_golden.py from long-double arithmetic that shares no
    code with the oracle's C kernels): values, Jacobians and fitted coefficients."""
    g = _gold()
    c, lb, ub, pts = (g[f"{name}_{k}"] for k in ("coefs", "lb", "ub", "pts"))
    N = len(lb)
    p = orc.OracleChebPoly(c, lb, ub)
    v, J = p.jacobian(pts if N > 1 else pts[:, 0])
    tol = 16 * np.finfo(np.float64).eps * np.abs(c).sum()
    nmax2 = max(1, max(n - 1 for n in c.shape[:N]) ** 2)
    assert np.abs(np.asarray(v).reshape(g[f"{name}_val"].shape) - g[f"{name}_val"]).max() <= tol
    assert np.abs(np.asarray(J).reshape(g[f"{name}_jac"].shape) - g[f"{name}_jac"]).max() <= tol * nmax2 * (2 / (ub - lb)).max()
    fit = orc.chebcoefs_fft(g[f"{name}_fitvals"], N)            # the FFT stand-in against the long-double cosine sums
    assert np.abs(fit - g[f"{name}_fitcoefs"]).max() <= 1e-14 * np.abs(g[f"{name}_fitcoefs"]).max()


def test_simd_baseline_variant_is_bit_identical(orc):
    """The component-vectorised evaluation used for the timed CPU baseline == the line-faithful restatement."""
    rng = np.random.default_rng(9)
    for dims, K, cplx in (((33, 33, 33), 3, False), ((9, 8), None, False), ((5, 4, 3, 6), 2, True), ((7,), 2, False),
                          ((2, 1, 3), 3, True), ((6, 5), 5, False)):
        shape = dims + ((K,) if K else ())
        c = rng.uniform(-1, 1, shape) + (1j * rng.uniform(-1, 1, shape) if cplx else 0)
        N = len(dims)
        p = orc.OracleChebPoly(c, -np.ones(N), np.ones(N))
        x = rng.uniform(-1, 1, (257, N))
        x = x if N > 1 else x[:, 0]
        assert np.array_equal(p(x), p.eval_simd(x))

"""GPU: SURVEY.md §8f N1/N3 — roots / colleague_matrix / chebinterp(f, ...) built on the device path
(fastcheb_reinterp: Chebyshev grid -> batched evaluation -> DCT-I fit -> droptol on the GPU; only the <= 50x50
eigenproblems run on the host, as in the reference).  Known answers: test/runtests.jl:28-34, README.md:63-84."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))


def test_roots_reference_testset(fc):
    assert len(fc.roots(fc.chebinterp(lambda x: 1.1, 0, 1))) == 0                      # runtests.jl:28
    assert np.allclose(fc.roots(fc.chebinterp(lambda x: x - 0.1, 0, 1)), [0.1])       # :29
    p = fc.chebinterp(lambda x: (x - 0.1) * (x - 0.2), 0, 1)
    assert p.coefs.size == GOLD["colleague"]["ncoefs"]                                 # :31 degree-2 polynomial
    assert np.allclose(fc.roots(p), [0.1, 0.2])                                        # :32
    assert np.allclose(fc.colleague_matrix(p), GOLD["colleague"]["colleague_matrix"])  # :34
    c = fc.chebinterp(np.cos, 10**4, 0, 1000)                                          # :33 (order-10^4 fit, n = 10001)
    r = fc.roots(c)
    assert len(r) == 318
    assert np.allclose(r / np.pi, np.arange(318) + 0.5, rtol=np.sqrt(np.finfo(np.float64).eps))


def test_reinterp_matches_host_grid_path(fc, orc):
    """chebinterp(c, order, lb, ub) on the device == fitting c's values on the host-generated grid (oracle fit)."""
    rng = np.random.default_rng(31)
    coefs = rng.uniform(-1, 1, (20, 15)) * 0.5 ** np.add.outer(np.arange(20), np.arange(15))
    lb, ub = np.array([0.0, -1.0]), np.array([2.0, 1.0])
    c = fc.ChebPoly(coefs, lb, ub)
    nlb, nub = np.array([0.5, -0.25]), np.array([1.5, 0.75])
    d = fc.chebinterp(c, (24, 20), nlb, nub, tol=0.0)
    x = orc.chebpoints((24, 20), nlb, nub)
    ref = orc.chebinterp(orc.OracleChebPoly(coefs, lb, ub)(x.reshape(-1, 2)).reshape(25, 21), nlb, nub, tol=0.0)
    assert d.coefs.shape == (25, 21)
    assert np.abs(d.coefs - ref.coefs).max() <= 1e-13 * np.abs(ref.coefs).max()
    pts = nlb + (nub - nlb) * rng.random((1000, 2))
    assert np.abs(d(pts) - c(pts)).max() <= 1e-13 * np.abs(coefs).sum()
    with pytest.raises(fc.ArgumentError):
        fc.chebinterp(c, (8, 8), [-1.0, -1.0], [1.0, 1.0])                            # new box leaves the domain


def test_adaptive_chebinterp(fc):
    """interp.jl:174-203 and runtests.jl:44-49: order doubling until converged; max_order is respected."""
    f = lambda x: np.exp(x[0] + 2 * x[1]) / (1 + 2 * x[0] ** 2 + x[1] ** 2)           # noqa: E731  (runtests.jl:41)
    lb, ub = [-0.3, 0.1], [0.9, 1.2]
    c = fc.chebinterp(f, lb, ub)
    assert all(n - 1 < m for n, m in zip(c.dims, (40, 30)))                           # runtests.jl:48
    x1 = np.array([0.2, 0.8])
    assert abs(c(x1) - f(x1)) <= 1e-13 * abs(f(x1)) * 50
    c2 = fc.chebinterp(f, lb, ub, max_order=(4, 5))
    assert tuple(n - 1 for n in c2.dims) == (4, 5)                                    # runtests.jl:49
    g = fc.chebinterp(lambda x: np.sin(2 * x + 3 * np.cos(4 * x)), 0, 10)             # README function, adaptive 1-D
    xs = np.linspace(0, 10, 101)
    assert np.abs(g(xs) - np.sin(2 * xs + 3 * np.cos(4 * xs))).max() < 1e-12
    with pytest.raises(fc.ArgumentError):
        fc.chebinterp(f, lb, ub, tol=0.0)

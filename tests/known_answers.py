"""The reference's own known answers for the hot path (README.md, test/runtests.jl), written once against
a tiny adapter so the SAME checks pin the CPU oracle (tests/test_oracle_golden.py, no GPU) and the CUDA
path (tests/test_gpu_golden.py, -m gpu).  Each check cites the reference lines it restates."""
import json
import os

import numpy as np

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))


class Api:
    """adapter: chebpoints(order, lb, ub, dtype), interp(vals, lb, ub, tol), value(c, x), grad(c, x), jac(c, x),
    interp_v1(vals, lb, ub), coefs(c)"""

    def __init__(self, chebpoints, interp, value, grad, jac, interp_v1, coefs, ArgumentError=Exception,
                 DomainError=Exception, DimensionMismatch=Exception):
        self.chebpoints, self.interp, self.value, self.grad, self.jac = chebpoints, interp, value, grad, jac
        self.interp_v1, self.coefs = interp_v1, coefs
        self.ArgumentError, self.DomainError, self.DimensionMismatch = ArgumentError, DomainError, DimensionMismatch


def rtol_default(T):
    return float(np.sqrt(np.finfo(T).eps))  # Julia's isapprox default


def check_readme_1d(api: Api):
    g = GOLD["readme_1d"]  # README.md:65-84
    f = lambda x: np.sin(2 * x + 3 * np.cos(4 * x))
    x = api.chebpoints(200, 0.0, 10.0, np.float64)
    c = api.interp(f(x), 0.0, 10.0, None)
    xx = np.arange(0, 101) * 0.1
    xx[-1] = 10.0
    err = np.max(np.abs(api.value(c, xx) - f(xx)))
    assert abs(err - g["max_err_on_0:0.1:10"]) < 1e-12, err
    v, dv = api.grad(c, 1.234)
    assert abs(v - g["chebgradient"][0]) < 1e-13
    assert abs(dv - g["chebgradient"][1]) < 1e-10


def check_readme_2d(api: Api):
    g = GOLD["readme_2d"]  # README.md:106-146
    gf = lambda x: np.sin(x[..., 0] + np.cos(x[..., 1]))
    lb, ub = [1, 3], [2, 4]
    x = api.chebpoints((10, 20), lb, ub, np.float64)
    c = api.interp(gf(x), lb, ub, None)
    assert tuple(n - 1 for n in api.coefs(c).shape) == tuple(g["order_default_tol"])  # README.md:131-132
    assert tuple(n - 1 for n in api.coefs(api.interp(gf(x), lb, ub, 0.0)).shape) == tuple(g["order_tol0"])
    assert tuple(n - 1 for n in api.coefs(api.interp(gf(x), lb, ub, 0.01)).shape) == tuple(g["order_tol0.01"])
    assert abs(api.value(c, np.array(g["x"])) - g["value"]) < 2e-15  # README.md:116-117
    v, gr = api.grad(c, np.array(g["x"]))
    assert abs(v - g["value"]) < 2e-15
    assert np.allclose(gr, g["gradient"], rtol=0, atol=1e-13)  # README.md:145-146


def check_1d_testset(api: Api, T):
    # runtests.jl:8-27
    lb, ub = T(-0.3), T(0.9)
    f = lambda x: np.exp(x) / (1 + 2 * x ** 2)
    fp = lambda x: f(x) * (1 - 4 * x / (1 + 2 * x ** 2))
    x = api.chebpoints(48, lb, ub, T)
    assert x.dtype == T
    c = api.interp(f(x).astype(T), lb, ub, 0.0)
    assert api.coefs(c).shape == (49,) and api.coefs(c).dtype == T
    x1 = T(0.2)
    rt = rtol_default(T)
    assert abs(api.value(c, x1) - f(0.2)) <= rt * abs(f(0.2))
    v, dv = api.grad(c, x1)
    assert abs(v - f(0.2)) <= rt * abs(f(0.2)) and abs(dv - fp(0.2)) <= rt * abs(fp(0.2))
    # runtests.jl:30-34: (x-0.1)(x-0.2) on [0,1] trims to exactly 3 coefficients
    xs = api.chebpoints(8, 0.0, 1.0, np.float64)
    p = api.interp((xs - 0.1) * (xs - 0.2), 0.0, 1.0, None)
    cf = api.coefs(p)
    assert cf.shape == (3,)
    # colleague matrix [0.5 -0.24; 0.5 -0.2] (runtests.jl:34; roots.jl:13-46 for n = 3) pins the coefficients
    assert np.allclose(cf, [0.245, 0.35, 0.125], rtol=0, atol=1e-15)
    C = np.array([[0.0, 0.5], [1.0, 0.0]])
    C[:, 1] -= cf[:2] / (2 * cf[2])
    C = C * 0.5 + 0.5 * np.eye(2)
    assert np.allclose(C, GOLD["colleague"]["colleague_matrix"], rtol=0, atol=1e-13)


def check_2d_testset(api: Api, T):
    # runtests.jl:38-86
    lb, ub = np.array([-0.3, 0.1], dtype=T), np.array([0.9, 1.2], dtype=T)
    den = lambda x: 1 + 2 * x[..., 0] ** 2 + x[..., 1] ** 2
    f = lambda x: np.exp(x[..., 0] + 2 * x[..., 1]) / den(x)
    gradf = lambda x: f(x)[..., None] * np.stack([1 - 4 * x[..., 0] / den(x), 2 - 2 * x[..., 1] / den(x)], -1)
    x = api.chebpoints((48, 39), lb, ub, T)
    assert x.dtype == T
    c = api.interp(f(x).astype(T), lb, ub, None)
    c0 = api.interp(f(x).astype(T), lb, ub, 0.0)
    assert api.coefs(c0).shape == (49, 40)
    x1 = np.array([0.2, 0.8], dtype=T)
    x1d = x1.astype(np.float64)
    rt = rtol_default(T)
    assert abs(api.value(c, x1) - f(x1d)) <= rt * abs(f(x1d))                       # :56
    assert abs(api.value(c, x1) - api.value(c0, x1)) <= 10 * np.finfo(T).eps * abs(api.value(c0, x1))  # :57
    assert all(a < b for a, b in zip(api.coefs(c).shape, api.coefs(c0).shape))      # :58
    v, g = api.grad(c, x1)
    assert abs(v - f(x1d)) <= rt * abs(f(x1d)) and np.allclose(g, gradf(x1d), rtol=rt, atol=0)  # :59
    # :63-67 univariate function in 2-D collapses dim 2 to size 1
    f1 = lambda x: np.exp(x[..., 0]) / (1 + 2 * x[..., 0] ** 2)
    c1 = api.interp(f1(x).astype(T), lb, ub, None)
    assert abs(api.value(c1, x1) - f1(x1d)) <= rt * abs(f1(x1d))
    assert api.coefs(c1).shape[1] == 1
    # :69-76 complex vector-valued
    CT = np.complex64 if T == np.float32 else np.complex128
    cis = lambda x: np.exp(1j * (x[..., 0] * x[..., 1] + 2 * x[..., 1]))
    f2 = lambda x: np.stack([f(x).astype(np.complex128), cis(x)], -1)
    jac2 = lambda x: np.stack([gradf(x).astype(np.complex128),
                               np.stack([1j * x[..., 1], 1j * (x[..., 0] + 2)], -1) * cis(x)[..., None]], -2)
    c2 = api.interp(f2(x).astype(CT), lb, ub, None)
    assert np.allclose(api.value(c2, x1), f2(x1d), rtol=rt, atol=0)
    v2, J2 = api.jac(c2, x1)
    assert np.allclose(v2, f2(x1d), rtol=rt, atol=0)
    assert np.allclose(J2, jac2(x1d), rtol=rt, atol=rt * 1e-2)
    # :78-84 chebinterp_v1: leading axis is the component axis
    av1 = np.moveaxis(f2(x).astype(CT), -1, 0)
    c2v1 = api.interp_v1(av1, lb, ub)
    assert np.allclose(api.value(c2v1, x1), f2(x1d), rtol=rt, atol=0)
    v3, J3 = api.jac(c2v1, x1)
    assert np.allclose(J3, jac2(x1d), rtol=rt, atol=rt * 1e-2)
    assert np.array_equal(api.coefs(c2v1), api.coefs(c2))


def check_degree1(api: Api):
    # runtests.jl:130-145
    lb, ub = -0.3, 0.9
    x = api.chebpoints(1, lb, ub, np.float64)
    c = api.interp(2 * x + 3, lb, ub, None)
    assert np.isclose(api.value(c, 0.0), 3) and np.isclose(api.value(c, 0.1), 3.2)
    assert np.isclose(api.grad(c, 0.0)[1], 2) and np.isclose(api.grad(c, 0.1)[1], 2)
    lb, ub = [-0.3, -0.1], [0.9, 1.2]
    x = api.chebpoints((1, 1), lb, ub, np.float64)
    c = api.interp((2 * x[..., 0] + 3) * (4 * x[..., 1] + 5), lb, ub, None)
    assert np.isclose(api.value(c, np.array([0.0, 0.0])), 15)
    assert np.isclose(api.value(c, np.array([0.1, 0.0])), 3.2 * 5)
    assert np.isclose(api.value(c, np.array([0.0, 0.1])), 3 * 5.4)
    assert np.allclose(api.grad(c, np.array([0.0, 0.0]))[1], [10, 12])


def check_degree0(api: Api):
    # runtests.jl:147-166 (gradients are exactly zero: ==, not ≈)
    lb, ub = -0.3, 0.9
    x = api.chebpoints(0, lb, ub, np.float64)
    c = api.interp(np.sin(x), lb, ub, None)
    assert np.isclose(api.value(c, 0.0), np.sin((lb + ub) / 2)) and np.isclose(api.value(c, 0.1), np.sin((lb + ub) / 2))
    assert api.grad(c, 0.0)[1] == 0 and api.grad(c, 0.1)[1] == 0
    lb, ub = [-0.3, -0.1], [0.9, 1.2]
    x = api.chebpoints((0, 0), lb, ub, np.float64)
    c = api.interp(np.sin(x[..., 0]) * np.cos(x[..., 1]), lb, ub, None)
    assert np.isclose(api.value(c, np.array([0.0, 0.0])), np.sin(0.3) * np.cos(0.55))
    assert np.isclose(api.value(c, np.array([0.1, 0.2])), np.sin(0.3) * np.cos(0.55))
    assert np.all(api.grad(c, np.array([0.1, 0.2]))[1] == 0)
    x = api.chebpoints((1, 0), lb, ub, np.float64)
    c = api.interp((2 * x[..., 0] + 3) * np.cos(x[..., 1]), lb, ub, None)
    assert np.isclose(api.value(c, np.array([0.0, 0.0])), 3 * np.cos(0.55))
    assert np.isclose(api.value(c, np.array([0.1, 0.22])), 3.2 * np.cos(0.55))
    g = api.grad(c, np.array([0.1, 0.2]))[1]
    assert np.isclose(g[0], 2 * np.cos(0.55)) and g[1] == 0


def check_int_bounds(api: Api):
    # runtests.jl:170: chebinterp(rand(5,5), (0,0), (1,1)) — integer bounds are accepted
    rng = np.random.default_rng(5)
    c = api.interp(rng.random((5, 5)), (0, 0), (1, 1), None)
    assert api.coefs(c).shape == (5, 5)
    assert np.isfinite(api.value(c, np.array([0.5, 0.25])))


def check_boundary(api: Api):
    # runtests.jl:174-179 (issue #36): x == ub is inside the domain; scalar and batched calls agree exactly
    x = api.chebpoints(50, 1.0, 100.0, np.float64)
    c = api.interp(np.sin(x), 1.0, 100.0, None)
    a = api.value(c, 100.0)
    b = api.value(c, np.array([100.0]))
    assert a == b[0]
    assert api.value(c, 1.0) == api.value(c, np.array([1.0, 50.0]))[0]


ALL_CHECKS = [check_readme_1d, check_readme_2d, check_degree1, check_degree0, check_int_bounds, check_boundary]
TYPED_CHECKS = [check_1d_testset, check_2d_testset]

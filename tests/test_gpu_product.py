"""The 1-D product-basis evaluation path (eval_pb1d.cuh, FASTCHEB_ALGO_PRODUCT): one multiply-add per coefficient.
Not bit-identical to the oracle (different summation order): values within the north_star 16 eps sum|c|, derivatives
within the differentiated-series bound, on point sets that include the faces and clusters next to them.  AUTO takes
the path only when its plan-time self-check against the Clenshaw kernel passes (coefficients that decay, i.e. fits of
smooth functions); dense-random coefficient vectors stay on the bit-exact Clenshaw kernel."""
import numpy as np
import pytest

from parity_util import jac_tol

pytestmark = pytest.mark.gpu

FUNCS = [
    ("readme", lambda x: np.sin(2 * x + 3 * np.cos(4 * x)), 0.0, 10.0),           # README.md:65-69 (configs[0])
    ("runge", lambda x: 1.0 / (1.0 + 25.0 * x * x), -1.0, 1.0),
    ("expcos", lambda x: np.exp(0.3 * x) * np.cos(7.0 * x + 0.2), -2.0, 3.5),
]


def _points(rng, lb, ub, npts, dt):
    x = (lb + (ub - lb) * rng.random(npts)).astype(dt)
    x[0], x[1] = lb, ub
    m = npts // 8
    d = (ub - lb) * 10.0 ** rng.uniform(-14, -2, m)
    x[2:2 + m] = np.where(rng.random(m) < 0.5, lb + d, ub - d).astype(dt)
    return np.clip(x, dt(lb), dt(ub))


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("n", [41, 65, 129, 201, 400, 512])
@pytest.mark.parametrize("fn", FUNCS, ids=[f[0] for f in FUNCS])
def test_product_path_matches_oracle_on_fits(fc, orc, fn, n, dt):
    name, f, lb, ub = fn
    rng = np.random.default_rng(n)
    xg = orc.chebpoints(n - 1, lb, ub)
    c = orc.chebcoefs(f(xg), 1).astype(dt)
    lbv, ubv = np.array([lb], dtype=dt), np.array([ub], dtype=dt)
    x = _points(rng, lb, ub, 30_011, dt)
    ref = orc.OracleChebPoly(c, lbv, ubv)
    v0, J0 = ref.jacobian(x)
    p = fc.ChebPoly(c, lbv, ubv)
    p.set_algo(fc.ALGO_PRODUCT, dt)
    assert p.get_algo(False, dt) == fc.ALGO_PRODUCT
    tol = 16 * np.finfo(dt).eps * np.abs(c).sum()
    v = p(x)
    assert v.dtype == dt and np.abs(v - v0).max() <= tol, np.abs(v - v0).max() / tol
    v2, J = fc.chebjacobian(p, x)
    assert np.array_equal(v2, v)
    assert np.abs(J[:, :, 0] - J0[:, :, 0]).max() <= jac_tol(c, 1, 0, lbv, ubv, dt)
    # the plan-time self-check sees the same thing, and AUTO then takes the path for large batches
    dv, dj = p.product_check(dt)
    assert dv <= 1.0 and dj <= 1.0, (dv, dj)
    p.set_algo(fc.ALGO_AUTO, dt)
    if dv <= 0.25:                                       # resolved fits (decaying coefficients): AUTO takes the path
        assert p.get_algo(False, dt) == fc.ALGO_PRODUCT
        assert p.get_algo(True, dt) == (fc.ALGO_PRODUCT if dj <= 0.25 else fc.ALGO_CLENSHAW)
        assert np.array_equal(p(x), v)                   # 30 011 points >= the planning threshold
        assert np.array_equal(p(x[:100]), v[:100])       # any batch size once planned
    else:                                                # under-resolved (e.g. the README function at n = 129): Clenshaw
        assert p.get_algo(False, dt) == fc.ALGO_CLENSHAW
        assert np.array_equal(p(x), v0)
    if name == "readme" and n == 201:
        assert dv <= 0.25 and dj <= 0.25                 # BASELINE configs[0] runs on the product-basis path


def test_auto_keeps_dense_random_coefficients_on_clenshaw(fc, orc):
    rng = np.random.default_rng(5)
    c = rng.uniform(-1, 1, 201)
    lb, ub = np.array([0.0]), np.array([10.0])
    x = _points(rng, 0.0, 10.0, 50_000, np.float64)
    p = fc.ChebPoly(c, lb, ub)
    ref = orc.OracleChebPoly(c, lb, ub)(x)
    assert np.array_equal(p(x), ref)                      # AUTO: the self-check rejects, Clenshaw stays (bit-identical)
    assert p.get_algo(False) == fc.ALGO_CLENSHAW
    dv, dj = p.product_check()
    assert dv > 0.25
    # forced, the path still evaluates the same polynomial to rounding: within tolerance away from the end points
    p.set_algo(fc.ALGO_PRODUCT)
    xi = 0.5 + 9.0 * rng.random(20_000)
    tol = 16 * np.finfo(float).eps * np.abs(c).sum()
    assert np.abs(p(xi) - orc.OracleChebPoly(c, lb, ub)(xi)).max() <= tol


@pytest.mark.parametrize("K,cplx", [(2, False), (3, False), (None, True), (2, True)], ids=["K2", "K3", "complex", "K2-complex"])
def test_product_path_vector_and_complex(fc, orc, K, cplx):
    rng = np.random.default_rng(17)
    n, lb, ub = 150, -1.0, 2.0
    xg = orc.chebpoints(n - 1, lb, ub)
    comps = []
    for k in range(K or 1):
        f = np.sin((2.0 + k) * xg + 0.3 * k) * np.exp(-0.2 * xg)
        if cplx:
            f = f + 1j * np.cos((1.5 + k) * xg)
        comps.append(f)
    vals = np.stack(comps, axis=-1) if K else comps[0]
    c = orc.chebcoefs(vals, 1)
    lbv, ubv = np.array([lb]), np.array([ub])
    x = _points(rng, lb, ub, 20_000, np.float64)
    v0, J0 = orc.OracleChebPoly(c, lbv, ubv).jacobian(x)
    p = fc.ChebPoly(c, lbv, ubv)
    p.set_algo(fc.ALGO_PRODUCT)
    tol = 16 * np.finfo(float).eps * np.abs(c).sum()
    v, J = fc.chebjacobian(p, x)
    assert np.abs(v - v0).max() <= tol and np.abs(p(x) - v0).max() <= tol
    assert np.abs(J - J0).max() <= jac_tol(c, 1, 0, lbv, ubv)
    # AD rules on this path (the Jacobian contracted in the kernel)
    dy = rng.uniform(-1, 1, v0.shape) + (1j * rng.uniform(-1, 1, v0.shape) if cplx else 0)
    y, pull = fc.rrule(p, x)
    want = np.real(np.einsum("pk,pk->p", np.conj(J0.reshape(len(x), -1)), dy.reshape(len(x), -1)))
    assert np.abs(np.asarray(pull(dy)).reshape(-1) - want).max() <= jac_tol(c, 1, 0, lbv, ubv) * (K or 1) * 2
    dx = rng.uniform(-1, 1, len(x))
    y2, ydot = fc.frule(dx, p, x)
    assert np.abs(np.asarray(ydot).reshape(len(x), -1) - J0.reshape(len(x), -1) * dx[:, None]).max() <= jac_tol(c, 1, 0, lbv, ubv)
    bad = x.copy()
    bad[12345] = ub + 0.5
    with pytest.raises(fc.ArgumentError) as ei:
        p(bad)
    assert ei.value.index == 12345


def test_product_path_unsupported_shapes(fc):
    rng = np.random.default_rng(3)
    for shape in ((30,), (600,), (65, 65), (65, 5)):
        lb, ub = -np.ones(len(shape) if len(shape) == 2 and shape[1] > 5 else 1), np.ones(len(shape) if len(shape) == 2 and shape[1] > 5 else 1)
        p = fc.ChebPoly(rng.uniform(-1, 1, shape), lb, ub)
        with pytest.raises(fc.CudaError):
            p.set_algo(fc.ALGO_PRODUCT)


def test_product_gradient_second_table_limit(fc, orc):
    """Gradient kernels carry a second coefficient table (the derivative series) in the kernel parameters; Float64 with 4
    real components and more than 496 coefficients does not fit: the forced path refuses gradients (values still work),
    AUTO keeps such gradients on the bit-exact Clenshaw kernel."""
    rng = np.random.default_rng(23)
    lbv, ubv = np.array([-1.0]), np.array([1.5])
    x = _points(rng, -1.0, 1.5, 20_000, np.float64)
    for n, covered in ((496, True), (500, False)):
        xg = orc.chebpoints(n - 1, -1.0, 1.5)
        c = orc.chebcoefs(np.stack([np.sin((40.0 + 9 * k) * xg + 0.3 * k) * np.exp(-0.2 * xg) for k in range(4)], axis=-1), 1)
        o = orc.OracleChebPoly(c, lbv, ubv)
        v0, J0 = o.jacobian(x)
        p = fc.ChebPoly(c, lbv, ubv)
        p.set_algo(fc.ALGO_PRODUCT)
        tol = 16 * np.finfo(float).eps * np.abs(c).sum(axis=0).max()
        assert np.abs(p(x) - v0).max() <= tol
        if covered:
            v, J = fc.chebjacobian(p, x)
            assert np.abs(v - v0).max() <= tol
            assert np.abs(J - J0).max() <= jac_tol(c, 1, 0, lbv, ubv)
        else:
            with pytest.raises(fc.CudaError):
                fc.chebjacobian(p, x)
            p.set_algo(fc.ALGO_AUTO)
            v, J = fc.chebjacobian(p, x)
            assert np.array_equal(v, v0) and np.array_equal(J, J0)


def test_product_path_refuses_spans_that_fail_the_division_test(fc, orc):
    """The product-basis kernel divides by the span with reciprocal + exact-remainder correction unconditionally; a span whose
    mantissa is all ones fails that test: forcing the path is refused and AUTO stays bit-identical to the oracle."""
    n, lb, ub = 101, 0.0, float(np.nextafter(2.0, 0.0))
    xg = orc.chebpoints(n - 1, lb, ub)
    c = orc.chebcoefs(np.sin(3 * xg) * np.exp(-xg), 1)
    lbv, ubv = np.array([lb]), np.array([ub])
    p = fc.ChebPoly(c, lbv, ubv)
    with pytest.raises(fc.CudaError):
        p.set_algo(fc.ALGO_PRODUCT)
    rng = np.random.default_rng(2)
    x = _points(rng, lb, ub, 20_000, np.float64)
    v0, J0 = orc.OracleChebPoly(c, lbv, ubv).jacobian(x)
    v, J = fc.chebjacobian(p, x)
    assert np.array_equal(v, v0) and np.array_equal(J, J0)

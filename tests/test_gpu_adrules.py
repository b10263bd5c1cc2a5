"""GPU: SURVEY.md §8f N4 — batched AD rules (chainrules.jl) with the Jacobian contracted inside the evaluation kernels
(fastcheb_eval_vjp / fastcheb_eval_jvp), both kernel families, vs the oracle's Jacobian contracted in numpy; plus the
reference's own finite-difference style checks (test_rrule / test_frule, runtests.jl:25-26, 75-76)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps

CASES = [((33, 33, 33), 3, False), ((17, 17, 17, 17), None, True), ((65, 65), None, False), ((9, 8, 7), 5, False),
         ((6, 5, 4), 4, True), ((201,), None, False), ((65,), 2, True), ((2, 1), None, False)]


@pytest.mark.parametrize("dims,K,cplx", CASES, ids=lambda v: str(v))
def test_vjp_jvp_match_oracle_jacobian(fc, orc, dims, K, cplx):
    rng = np.random.default_rng(91)
    shape = dims + ((K,) if K else ())
    c = rng.uniform(-1, 1, shape)
    if cplx:
        c = c + 1j * rng.uniform(-1, 1, shape)
    N = len(dims)
    lb, ub = -1.0 - 0.25 * np.arange(N), 0.5 + 0.75 * np.arange(N)
    npts = 3000
    pts = lb + (ub - lb) * rng.random((npts, N))
    pts[0], pts[1] = lb, ub
    ref = orc.OracleChebPoly(c, lb, ub)
    v0, J0 = ref.jacobian(pts if N > 1 else pts[:, 0])
    J0 = np.asarray(J0).reshape(npts, K or 1, N)
    dy = rng.uniform(-1, 1, (npts, K or 1)) + (1j * rng.uniform(-1, 1, (npts, K or 1)) if cplx else 0)
    dx = rng.uniform(-1, 1, (npts, N))
    want_xbar = np.real(np.einsum("pkn,pk->pn", np.conj(J0), dy))
    want_ydot = np.einsum("pkn,pn->pk", J0, dx)
    p = fc.ChebPoly(c, lb, ub)
    from parity_util import jac_tol
    # contractions of the Jacobian: the value bound plus the per-column differentiated-series bounds
    tol = 16 * EPS * np.abs(c).sum() + sum(jac_tol(c, N, d, lb, ub) for d in range(N))
    algos = [fc.ALGO_CLENSHAW] + ([fc.ALGO_DMMA] if (N >= 2 and dims[0] >= 5 and np.prod(dims[1:]) >= 8) else [])
    for algo in algos:
        p.set_algo(algo)
        xin = pts if N > 1 else pts[:, 0]
        y, pull = fc.rrule(p, xin)
        xbar = pull(dy if K else dy[:, 0])
        assert np.abs(np.asarray(y).reshape(npts, -1) - np.asarray(v0).reshape(npts, -1)).max() <= tol
        assert np.abs(np.asarray(xbar).reshape(npts, N) - want_xbar).max() <= tol * (K or 1) * 2, algo
        y2, ydot = fc.frule(dx if N > 1 else dx[:, 0], p, xin)
        assert np.abs(np.asarray(ydot).reshape(npts, -1) - want_ydot.reshape(npts, -1)).max() <= tol * N, algo


def test_rules_against_finite_differences(fc):
    """The reference's test_rrule/test_frule are finite-difference checks at rtol = atol = sqrt(eps) (runtests.jl:25-26)."""
    lb, ub = np.array([-0.3, 0.1]), np.array([0.9, 1.2])
    f = lambda x: np.exp(x[..., 0] + 2 * x[..., 1]) / (1 + 2 * x[..., 0] ** 2 + x[..., 1] ** 2)  # noqa: E731 runtests.jl:41
    c = fc.chebinterp(f(fc.chebpoints((48, 39), lb, ub)), lb, ub)
    x = np.array([0.2, 0.8])
    h = 1e-6
    fd = np.array([(c(x + h * e) - c(x - h * e)) / (2 * h) for e in np.eye(2)])
    y, pull = fc.rrule(c, x)
    assert np.allclose(pull(1.0), fd, rtol=1e-7, atol=1e-7) and np.isclose(y, f(x))
    dx = np.array([0.3, -0.7])
    y2, ydot = fc.frule(dx, c, x)
    assert np.isclose(ydot, fd @ dx, rtol=1e-7, atol=1e-7)
    # perturbing the polynomial itself (chainrules.jl:27-37): coefficients and bounds
    rng = np.random.default_rng(5)
    dco = rng.uniform(-1, 1, c.coefs.shape) * 1e-3
    dlb, dub = np.array([0.01, -0.02]), np.array([0.03, 0.015])
    t = 1e-6
    cp = fc.ChebPoly(c.coefs + t * dco, lb + t * dlb, ub + t * dub)
    cm = fc.ChebPoly(c.coefs - t * dco, lb - t * dlb, ub - t * dub)
    fd_all = (cp(x + t * dx) - cm(x - t * dx)) / (2 * t)
    _, ydot_all = fc.frule(dx, c, x, dself=fc.ChebPoly(dco, lb, ub), dlb=dlb, dub=dub)
    assert np.isclose(ydot_all, fd_all, rtol=1e-6, atol=1e-7)

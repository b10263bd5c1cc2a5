"""GPU: size-independent properties at BASELINE.json's full workload shapes (too many points for the CPU
oracle to check one by one): agreement of the two independent evaluation kernels (the Clenshaw kernel is
bit-identical to the oracle, tests/test_gpu_eval.py, so this carries oracle parity to millions of points),
partition invariance, linearity, and fit -> evaluate round trips."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def _dev_points(torch, n, N, lb, ub, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    u = torch.rand((n, N), dtype=torch.float64, device="cuda", generator=g)
    lbt, ubt = torch.tensor(lb, device="cuda"), torch.tensor(ub, device="cuda")
    return torch.minimum(torch.maximum(lbt + (ubt - lbt) * u, lbt), ubt)


def test_cfg3_kernels_agree_on_16M_points(fc):
    """configs[2]: 3-D order (32,32,32), SVector{3}: tensor-core path == Clenshaw path within the north_star
    tolerance on 2^24 device-resident points, value and Jacobian."""
    import torch

    rng = np.random.default_rng(4)
    coefs = rng.uniform(-1, 1, (33, 33, 33, 3))
    lb, ub = np.array([-1.0, -2.0, 0.5]), np.array([1.0, 3.0, 0.75])
    c = fc.ChebPoly(coefs, lb, ub)
    tol = 16 * EPS * np.abs(coefs).sum()
    npts = 1 << 24
    pts = _dev_points(torch, npts, 3, lb, ub, 5)
    c.set_algo(fc.ALGO_DMMA)
    v_t = c(pts)
    c.set_algo(fc.ALGO_CLENSHAW)
    v_c = c(pts)
    assert float((v_t - v_c).abs().max()) <= tol
    # partition invariance: a point's result does not depend on where it sits in the batch (bitwise)
    c.set_algo(fc.ALGO_DMMA)
    cut = 5_000_003
    assert torch.equal(c(pts[:cut]), v_t[:cut]) and torch.equal(c(pts[cut:]), v_t[cut:])
    # Jacobian on a 2^22 slice
    sl = pts[: 1 << 22]
    vt, Jt = fc.chebjacobian(c, sl)
    c.set_algo(fc.ALGO_CLENSHAW)
    vc, Jc = fc.chebjacobian(c, sl)
    from parity_util import jac_tol
    assert float((vt - vc).abs().max()) <= tol
    for d in range(3):   # the differentiated-series bound per Jacobian column (tests/parity_util.py)
        assert float((Jt[:, :, d] - Jc[:, :, d]).abs().max()) <= jac_tol(coefs, 3, d, lb, ub), d
    assert torch.equal(vt, v_t[: 1 << 22])       # the Jacobian kernel returns the same values as the value kernel


def test_cfg3_linearity(fc):
    import torch

    rng = np.random.default_rng(40)
    a, b = rng.uniform(-1, 1, (33, 33, 33, 3)), rng.uniform(-1, 1, (33, 33, 33, 3))
    lb, ub = -np.ones(3), np.ones(3)
    pts = _dev_points(torch, 1 << 22, 3, lb, ub, 41)
    ca, cb, cab = (fc.ChebPoly(x, lb, ub) for x in (a, b, 2.0 * a - 0.5 * b))
    tol = 16 * EPS * (2.0 * np.abs(a).sum() + 0.5 * np.abs(b).sum() + np.abs(2 * a - 0.5 * b).sum())
    assert float((cab(pts) - (2.0 * ca(pts) - 0.5 * cb(pts))).abs().max()) <= tol


def test_cfg4_4d_complex_slab_tiled(fc, orc):
    """configs[3]: 4-D order (16,16,16,16) ComplexF64 (1.34 MB of coefficients: slab-tiled through shared memory)."""
    import torch

    rng = np.random.default_rng(6)
    coefs = rng.uniform(-1, 1, (17,) * 4) + 1j * rng.uniform(-1, 1, (17,) * 4)
    lb, ub = -np.ones(4), np.array([1.0, 2.0, 3.0, 4.0])
    c = fc.ChebPoly(coefs, lb, ub)
    tol = 16 * EPS * np.abs(coefs).sum()
    pts = _dev_points(torch, 1 << 21, 4, lb, ub, 7)
    c.set_algo(fc.ALGO_DMMA)
    vt = c(pts)
    c.set_algo(fc.ALGO_CLENSHAW)
    vc = c(pts)
    assert float((vt - vc).abs().max()) <= tol
    sample = pts[:2000].cpu().numpy()
    want = orc.OracleChebPoly(coefs, lb, ub)(sample)
    assert np.array_equal(vc[:2000].cpu().numpy(), want)           # Clenshaw kernel: bit-identical to the oracle
    assert np.abs(vt[:2000].cpu().numpy() - want).max() <= tol


def test_cfg2_fit_then_evaluate_with_gradient(fc):
    """configs[1]: 2-D order (64,64) fit of runtests.jl:41's function, then 10^7 evaluations with chebgradient
    against the analytic function and gradient."""
    import torch

    lb, ub = np.array([-0.3, 0.1]), np.array([0.9, 1.2])
    f = lambda x: np.exp(x[..., 0] + 2 * x[..., 1]) / (1 + 2 * x[..., 0] ** 2 + x[..., 1] ** 2)  # noqa: E731
    x = fc.chebpoints((64, 64), lb, ub)
    c = fc.chebinterp(f(x), lb, ub, tol=0.0)
    assert c.coefs.shape == (65, 65)
    pts = _dev_points(torch, 10_000_000, 2, lb, ub, 3)
    v, g = fc.chebgradient(c, pts)
    p = pts.cpu().numpy()
    fx = f(p)
    den = 1 + 2 * p[:, 0] ** 2 + p[:, 1] ** 2
    gx = np.stack([fx * (1 - 4 * p[:, 0] / den), fx * (2 - 2 * p[:, 1] / den)], -1)
    assert np.abs(v.cpu().numpy() - fx).max() <= 1e-11 * np.abs(fx).max()
    assert np.abs(g.cpu().numpy() - gx).max() <= 1e-9 * np.abs(gx).max()


def test_cfg5_fit_then_evaluate_roundtrip(fc):
    """configs[4]: 10^5 independent order-64 fits of f_i(x) = sin(a_i x + b_i), then every polynomial evaluated at its
    own Chebyshev points reproduces the data (Float64 and Float32)."""
    rng = np.random.default_rng(8)
    M, n = 100_000, 65
    a, b = rng.uniform(0.5, 3.0, M), rng.uniform(-1, 1, M)
    grid = np.cos(np.pi * np.arange(n) / (n - 1))                   # interp.jl:6 on [-1, 1]
    vals = np.sin(a[:, None] * grid[None, :] + b[:, None])
    for dt, tol in ((np.float64, 1e-13), (np.float32, 450 * np.finfo(np.float32).eps)):
        C = fc.chebcoefs(vals.astype(dt), ndim=1, howmany=M)
        # evaluate sum_k C[i,k] T_k(x_j) for a subset of fits on the device path: one ChebPoly per sampled fit
        for i in rng.integers(0, M, 8):
            p = fc.ChebPoly(C[i], [-1.0], [1.0])
            got = p(grid.astype(dt))
            assert np.abs(got - vals[i]).max() <= 40 * tol, (dt, i)
        # Parseval-type checksum over the whole batch: c_0 equals the Clenshaw-Curtis mean of the data
        w = np.full(n, 1.0 / (n - 1))
        w[0] = w[-1] = 0.5 / (n - 1)
        assert np.abs(C[:, 0] - vals @ w).max() <= 10 * tol

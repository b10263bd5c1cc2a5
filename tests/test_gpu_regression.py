"""GPU: SURVEY.md §8f N2 — chebregression / chebvandermonde (regression.jl).  The Vandermonde matrix is assembled by
the CUDA kernel; known answers are the reference's own regression testsets (test/runtests.jl:88-128) and the README
example (README.md:86-101)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_1d_regression_testset(fc, orc):
    x = np.linspace(1, 2, 200)                                             # runtests.jl:89-93
    y = np.sin(x + 0.2) * x
    c = fc.chebregression(x, y, 2)
    c2 = np.linalg.lstsq(np.vander(x, 3, increasing=True), y, rcond=None)[0]
    assert c(1.1) == pytest.approx(c2[0] + 1.1 * c2[1] + 1.1 ** 2 * c2[2], rel=1e-13)
    xs = np.array([0.14, 0.95, 0.83, 0.13, 0.42, 0.12])                   # runtests.jl:96-102
    A1 = fc.chebvandermonde(xs, 0.1, 0.99, 4)
    # the reference's generic routine: column i = the interpolant with unit coefficient e_i evaluated at x
    A = np.stack([orc.OracleChebPoly(np.eye(5)[i], np.array([0.1]), np.array([0.99]))(xs) for i in range(5)], axis=1)
    assert A1.shape == (6, 5) and np.allclose(A1, A, rtol=1e-13, atol=1e-15)
    with pytest.raises(fc.ArgumentError):
        fc.chebvandermonde(xs, 0.1, 0.9, 4)                               # 0.95 outside the domain
    # 1-D recurrence is the reference's (regression.jl:35-43): T_i = 2x*T_{i-1} - T_{i-2} with separate roundings
    z = (xs - 0.1) * 2 / (0.99 - 0.1) - 1
    T = [np.ones_like(z), z]
    for _ in range(3):
        T.append(2 * z * T[-1] - T[-2])
    assert np.array_equal(A1, np.stack(T, axis=1))


def test_2d_regression_testset(fc):
    X1, X2 = np.meshgrid(np.linspace(-3, 2, 50), np.linspace(0, 1, 40), indexing="ij")   # runtests.jl:106
    x = np.stack([X1.ravel(order="F"), X2.ravel(order="F")], axis=1)
    fr = lambda p: np.sin(p[..., 0] / 3 + 1) * np.exp(p[..., 1])          # noqa: E731
    c = fc.chebregression(x, fr(x), (2, 3))
    pows = lambda p: np.stack([p[..., 0] ** i * p[..., 1] ** j for j in range(4) for i in range(3)], axis=-1)  # noqa: E731
    c2 = np.linalg.lstsq(pows(x), fr(x), rcond=None)[0]
    p = np.array([1.12345, np.sqrt(0.5)])
    assert c(p) == pytest.approx(pows(p) @ c2, rel=1e-12)                 # runtests.jl:115
    fr4 = lambda q: np.cos(q[..., 0] / 3 + 1) * np.sin(q[..., 1])         # noqa: E731
    c4 = fc.chebregression(x, fr4(x), (2, 3))
    c14 = fc.chebregression(x, np.stack([fr(x), fr4(x)], axis=1), (2, 3))  # vector-valued (:120-124)
    assert np.allclose(c14.coefs[..., 0], c.coefs, rtol=1e-12, atol=1e-14)
    assert np.allclose(c14.coefs[..., 1], c4.coefs, rtol=1e-12, atol=1e-14)
    with pytest.raises(fc.ArgumentError):
        fc.chebregression(x[:5], fr(x[:5]), (2, 3))                       # not enough points (regression.jl:55)


def test_readme_regression_and_device_path(fc):
    """README.md:86-101: 10^4 random points, degree 200 — the fit is as good as the interpolant (max error ~1.5e-5);
    the device-resident path (torch in, cuSOLVER QR) gives the same polynomial."""
    import torch

    f = lambda x: np.sin(2 * x + 3 * np.cos(4 * x))                        # noqa: E731
    rng = np.random.default_rng(0)
    xr = rng.random(10**4) * 10
    c = fc.chebregression(xr, f(xr), 0, 10, 200)
    xx = np.arange(0, 10.05, 0.1)
    err = np.abs(c(np.clip(xx, 0, 10)) - f(xx)).max()
    assert 1e-6 < err < 1e-4                                               # README: 1.4655e-5 for its random draw
    cd = fc.chebregression(torch.as_tensor(xr, device="cuda"), torch.as_tensor(f(xr), device="cuda"), 0, 10, 200)
    assert np.abs(cd.coefs - c.coefs).max() <= 1e-7 * np.abs(c.coefs).max()   # two QR implementations, cond(A) ~ 1e3
    A = fc.chebvandermonde(torch.as_tensor(rng.random((1000, 3)), device="cuda"), [0, 0, 0], [1, 1, 1], (3, 2, 4))
    assert tuple(A.shape) == (1000, 60) and A.stride() == (1, 1000)
    Ah = fc.chebvandermonde(A.new_tensor(rng.random((50, 3))).cpu().numpy(), [0, 0, 0], [1, 1, 1], (3, 2, 4))
    assert Ah.shape == (50, 60)

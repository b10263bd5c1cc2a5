"""Low-order 1-D evaluation path (eval_lo1d.cuh; n <= 17, batches >= 4096 points, under AUTO): vector loads / stores,
coefficients in the kernel parameters.  Bit-identical to the CPU oracle (eval.jl:19-31, :88-101, :151) and to the generic
Clenshaw kernel for every n (including the n = 1, 2 branches and the zero-padded recurrence steps), ragged batch tails,
out-of-domain points inside a vector group, and pointers that are not 16-byte aligned (generic kernel takes over)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _poly(fc, rng, n, K, cplx, dt):
    shape = (n,) + ((K,) if K else ())
    c = rng.uniform(-1, 1, shape)
    if cplx:
        c = c + 1j * rng.uniform(-1, 1, shape)
        c = c.astype(np.complex128 if dt == np.float64 else np.complex64)
    else:
        c = c.astype(dt)
    lb, ub = np.array([-0.7], dtype=dt), np.array([2.1], dtype=dt)
    return c, lb, ub


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("K,cplx", [(None, False), (2, False), (3, False), (None, True), (2, True), (4, False)],
                         ids=["scalar", "K2", "K3", "complex", "K2-complex", "K4"])
def test_low_order_bit_identical(fc, orc, dt, K, cplx):
    rng = np.random.default_rng(41)
    for n in (1, 2, 3, 4, 5, 6, 8, 9, 10, 13, 16, 17, 18):   # 18: past the path (generic kernel), same assertions
        c, lb, ub = _poly(fc, rng, n, K, cplx, dt)
        npts = int(rng.integers(4096, 9000)) | 1              # ragged tail: not a multiple of the points per thread
        x = (lb[0] + (ub[0] - lb[0]) * rng.random(npts)).astype(dt)
        x[0], x[1], x[-1] = lb[0], ub[0], ub[0]
        x = np.clip(x, lb[0], ub[0])
        o = orc.OracleChebPoly(c, lb, ub)
        v0, J0 = o.jacobian(x)
        p = fc.ChebPoly(c, lb, ub)
        v, J = fc.chebjacobian(p, x)
        assert np.array_equal(p(x), v0), (n, "value")
        assert np.array_equal(v, v0) and np.array_equal(J, J0), (n, "jacobian")
        p.set_algo(fc.ALGO_CLENSHAW, dt)                       # the generic kernel
        v2, J2 = fc.chebjacobian(p, x)
        assert np.array_equal(v2, v0) and np.array_equal(J2, J0)


def test_low_order_domain_errors_and_dead_points(fc):
    rng = np.random.default_rng(5)
    c = rng.uniform(-1, 1, 7)
    p = fc.ChebPoly(c, [0.0], [1.0])
    x = rng.random(50_001)
    good = p(x)
    for bad_at, val in ((0, -1e-9), (4097, 1.0 + 1e-12), (50_000, np.nan), (12_346, 7.0)):
        bad = x.copy()
        bad[bad_at] = val
        bad[min(bad_at + 9, len(bad) - 1)] = 3.0 if bad_at + 9 < len(bad) else bad[-1]   # a later offender: the FIRST index is reported
        with pytest.raises(fc.ArgumentError) as ei:
            p(bad)
        assert ei.value.index == bad_at
        with pytest.raises(fc.ArgumentError) as ei:
            fc.chebgradient(p, bad)
        assert ei.value.index == bad_at
    assert np.array_equal(p(x), good)


def test_low_order_device_pointers_and_misalignment(fc, orc):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(8)
    for dt, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
        c = rng.uniform(-1, 1, 9).astype(dt)
        lb, ub = np.array([-1.0], dtype=dt), np.array([1.0], dtype=dt)
        o = orc.OracleChebPoly(c, lb, ub)
        p = fc.ChebPoly(c, lb, ub)
        x = np.clip((2 * rng.random(20_011) - 1).astype(dt), -1, 1)
        xd = torch.from_numpy(x).cuda()
        for off in (0, 1, 2, 3, 5):       # element offsets: most are not 16-byte aligned -> generic kernel, same bits
            v0, J0 = o.jacobian(x[off:])
            v, J = fc.chebjacobian(p, xd[off:])
            assert np.array_equal(v.cpu().numpy(), v0) and np.array_equal(J.cpu().numpy().reshape(J0.shape), J0), (dt, off)
            assert np.array_equal(p(xd[off:]).cpu().numpy(), v0)


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_low_order_division_corner_cases(fc, orc, dt):
    """The path divides by the span with a reciprocal + exact-remainder correction: spans whose mantissa is all ones stay on
    the generic kernel, and derivative numerators near the overflow / underflow thresholds take the true division."""
    rng = np.random.default_rng(77)
    big, small = (1e200, 1e-200) if dt == np.float64 else (1e30, 1e-30)
    cases = [(0.0, float(np.nextafter(dt(2.0), dt(0.0))), 1.0), (-1.0, 1.0, big), (-1.0, 1.0, small), (0.25, 0.75, small * 1e-5)]
    for lb0, ub0, scale in cases:
        c = (rng.uniform(-1, 1, 11) * scale).astype(dt)
        c[3] = 0
        lb, ub = np.array([lb0], dtype=dt), np.array([ub0], dtype=dt)
        x = np.clip((lb0 + (ub0 - lb0) * rng.random(8192)).astype(dt), lb[0], ub[0])
        x[5], x[6] = lb[0], ub[0]
        with np.errstate(all="ignore"):
            v0, J0 = orc.OracleChebPoly(c, lb, ub).jacobian(x)
        v, J = fc.chebjacobian(fc.ChebPoly(c, lb, ub), x)
        assert np.array_equal(v, v0, equal_nan=True) and np.array_equal(J, J0, equal_nan=True), (lb0, ub0, scale)


ND_SHAPES = [(5, 5), (2, 7), (9, 1), (1, 6), (9, 9), (4, 4, 4), (3, 1, 5), (2, 2, 2), (5, 5, 5), (3, 3, 3, 3), (2, 4, 3, 2), (13, 3),
             (9, 9, 9), (9, 30), (5, 5, 5, 5), (4, 60), (7, 7, 7, 2)]   # the last four: second table capacity (up to 2048 padded coefficients)


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("K,cplx", [(None, False), (2, False), (3, False), (None, True), (4, False)],
                         ids=["scalar", "K2", "K3", "complex", "K4"])
def test_low_order_nd_bit_identical(fc, orc, dt, K, cplx):
    """N-d low-order path (eval_lond.cuh): bit-identical to the oracle and the generic kernel, ragged tails, face points,
    size-1 dimensions; shapes the path does not take (register budget, coefficient capacity) run the same assertions."""
    rng = np.random.default_rng(43)
    for shape in ND_SHAPES:
        N = len(shape)
        full = shape + ((K,) if K else ())
        c = rng.uniform(-1, 1, full)
        if cplx:
            c = (c + 1j * rng.uniform(-1, 1, full)).astype(np.complex128 if dt == np.float64 else np.complex64)
        else:
            c = c.astype(dt)
        lb = rng.uniform(-2, -0.5, N).astype(dt)
        ub = rng.uniform(0.5, 3, N).astype(dt)
        npts = int(rng.integers(4096, 7000)) | 1
        x = (lb + (ub - lb) * rng.random((npts, N))).astype(dt)
        x[0], x[1], x[-1] = lb, ub, ub
        x[2, 0] = lb[0]
        x = np.clip(x, lb, ub)
        o = orc.OracleChebPoly(c, lb, ub)
        v0, J0 = o.jacobian(x)
        p = fc.ChebPoly(c, lb, ub)
        if p.get_algo(False, dt) == fc.ALGO_DMMA:   # AUTO sends this shape to the tensor cores (tolerance parity, not bits)
            continue
        v, J = fc.chebjacobian(p, x)
        assert np.array_equal(p(x), v0), (shape, "value")
        assert np.array_equal(v, v0) and np.array_equal(J, J0), (shape, "jacobian")
        p.set_algo(fc.ALGO_CLENSHAW, dt)
        v2, J2 = fc.chebjacobian(p, x)
        assert np.array_equal(v2, v0) and np.array_equal(J2, J0)


def test_low_order_nd_domain_errors(fc):
    rng = np.random.default_rng(6)
    p = fc.ChebPoly(rng.uniform(-1, 1, (4, 5, 3)), [0.0, 0.0, -1.0], [1.0, 2.0, 1.0])
    x = rng.random((30_001, 3)) * [1, 2, 2] - [0, 0, 1]
    good = p(x)
    for bad_at, d, val in ((0, 2, -1.0000001), (4099, 1, 2.0 + 1e-9), (30_000, 0, np.nan), (12_347, 1, -0.5)):
        bad = x.copy()
        bad[bad_at, d] = val
        if bad_at + 3 < len(bad):
            bad[bad_at + 3, 0] = 9.0      # a later offender: the FIRST index is reported
        with pytest.raises(fc.ArgumentError) as ei:
            p(bad)
        assert ei.value.index == bad_at
        with pytest.raises(fc.ArgumentError) as ei:
            fc.chebgradient(p, bad)
        assert ei.value.index == bad_at
    assert np.array_equal(p(x), good)


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("K", [None, 3], ids=["scalar", "K3"])
def test_lo1d_is_the_clenshaw_kernel_for_longer_series(fc, orc, dt, K):
    """Up to 513 coefficients the same kernel is AUTO's Clenshaw kernel for every 1-D polynomial the product-basis path does
    not take (n < 40, or rejected by its self-check — dense random coefficients are): bit-identical to the oracle, table
    capacities 17 / 129 / 513 and the generic kernel beyond."""
    rng = np.random.default_rng(47)
    for n in (18, 25, 39, 40, 64, 65, 128, 129, 130, 257, 512, 513, 514, 700):
        c = rng.uniform(-1, 1, (n,) + ((K,) if K else ())).astype(dt)
        lb, ub = np.array([-1.5], dtype=dt), np.array([0.5], dtype=dt)
        npts = 8192 + 3
        x = np.clip((lb[0] + (ub[0] - lb[0]) * rng.random(npts)).astype(dt), lb[0], ub[0])
        x[0], x[-1] = lb[0], ub[0]
        p = fc.ChebPoly(c, lb, ub)
        v, J = fc.chebjacobian(p, x)
        assert p.get_algo(False, dt) != fc.ALGO_PRODUCT and p.get_algo(True, dt) != fc.ALGO_PRODUCT, n
        v0, J0 = orc.OracleChebPoly(c, lb, ub).jacobian(x)
        assert np.array_equal(p(x), v0), (n, "value")
        assert np.array_equal(v, v0) and np.array_equal(J, J0), (n, "jacobian")

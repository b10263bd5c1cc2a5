"""GPU parity: batched evaluation / Jacobian through the C ABI vs the CPU oracle on the same seeded inputs.

Tolerances (north_star): |dv| <= 16 eps sum|c| per evaluated value.  The Clenshaw kernels execute the
oracle's exact IEEE operation sequence, so they are additionally required to be BIT-IDENTICAL to it.
Jacobian entries of the tensor-core (basis-form) path are held to the same bound applied to the
differentiated series, 16 eps (sum_k |c_k| k_d^2) * 2/(ub_d-lb_d) (tests/parity_util.py; measured worst case 0.18 of
it on BASELINE configs[1..3], profiles/r02_jac_parity.jsonl)."""
import numpy as np
import pytest

from parity_util import jac_tol

pytestmark = pytest.mark.gpu

CASES = [
    # dims, K, complex, dtype
    ((201,), None, False, np.float64),
    ((65,), None, False, np.float32),
    ((65,), 2, True, np.float64),
    ((65, 65), None, False, np.float64),
    ((49, 40), None, False, np.float32),
    ((33, 33, 33), 3, False, np.float64),
    ((33, 33, 33), 3, False, np.float32),
    ((17, 9, 5, 4), 2, True, np.float32),
    ((17, 17, 17, 17), None, True, np.float64),
    ((9, 8, 7), 5, False, np.float64),       # E = 5: two component groups
    ((6, 5, 4), 4, True, np.float32),        # E = 8
    ((1,), None, False, np.float64),         # n == 1 (eval.jl:22)
    ((2,), None, False, np.float64),         # n == 2 (eval.jl:23)
    ((2, 1), None, False, np.float64),       # mixed orders (1, 0) (runtests.jl:160-165)
    ((1, 1, 1), 2, False, np.float64),
    ((3, 1, 4), None, False, np.float64),
    ((34, 3), None, False, np.float64),
    ((5, 70), 3, False, np.float64),
    # 2-D shapes of the shared-memory-resident tensor-core kernel (eval_dmma2d.cuh): trailing columns k2 = 8j+1 / 8j+2
    # contracted on the CUDA cores (from the tail slots and from the main weight table), every KS, EC = 1..3, two groups
    ((33, 33), 3, False, np.float64),
    ((17, 10), None, True, np.float32),
    ((65, 66), None, False, np.float64),
    ((65, 67), None, False, np.float64),         # n2 > 4 KS + 2: falls back to the ring kernel
    ((65, 33), 2, False, np.float64),
    ((40, 16), None, False, np.float64),
    ((9, 9), 5, False, np.float64),
    ((30, 25), 2, True, np.float64),
    # 3-D / 4-D shapes of the ring kernel (eval_dmma.cuh) with n2 = 8 j + {0,1,2} columns per row of dimension 2
    ((33, 34, 5), None, False, np.float64),      # two trailing columns
    ((9, 16, 6, 3), 2, False, np.float64),       # no trailing column, 4-D
    ((65, 17, 4), None, True, np.float32),       # KS = 16
    ((12, 10, 7), 4, False, np.float64),         # two component groups
    ((41, 41), None, False, np.float64),         # 33 < n1 <= 49: 12 k-steps
    ((45, 9, 5), 2, False, np.float32),
    ((49, 50), None, True, np.float64),          # n2 = 4 KS + 2 at KS = 12
    ((4, 3, 5, 2, 3), 2, False, np.float64),     # 5-D and 6-D: the Clenshaw nest covers FASTCHEB_MAX_NDIM
    ((3, 2, 4, 2, 3, 2), None, True, np.float64),
    ((3, 3, 2, 2, 2, 3), 3, False, np.float32),
]


def make(dims, K, cplx, dt, seed):
    rng = np.random.default_rng(seed)
    shape = dims + ((K,) if K else ())
    c = rng.uniform(-1, 1, shape)
    if cplx:
        c = c + 1j * rng.uniform(-1, 1, shape)
        c = c.astype(np.complex64 if dt == np.float32 else np.complex128)
    else:
        c = c.astype(dt)
    N = len(dims)
    lb = (-1.0 - 0.25 * np.arange(N)).astype(dt)
    ub = (0.5 + 0.75 * np.arange(N)).astype(dt)
    return rng, c, lb, ub


def points(rng, lb, ub, npts, dt):
    N = len(lb)
    pts = (lb + (ub - lb) * rng.random((npts, N))).astype(dt)
    pts = np.clip(pts, lb, ub)
    pts[0], pts[1] = lb, ub                      # closed box: faces are inside (runtests.jl:174-179)
    if npts > 4:
        pts[2] = lb
        pts[2, -1] = ub[-1]
    return pts


def ids(case):
    dims, K, cplx, dt = case
    return f"{'x'.join(map(str, dims))}-K{K}-{'c' if cplx else 'r'}-{np.dtype(dt).name}"


@pytest.mark.parametrize("case", CASES, ids=ids)
def test_value_parity(fc, orc, case):
    dims, K, cplx, dt = case
    rng, c, lb, ub = make(*case, seed=11)
    pts = points(rng, lb, ub, 3001, dt)
    ref = orc.OracleChebPoly(c, lb, ub)(pts)
    poly = fc.ChebPoly(c, lb, ub)
    tol = 16 * np.finfo(dt).eps * np.abs(c).sum()
    poly.set_algo(fc.ALGO_CLENSHAW, dt)
    got = poly(pts)
    assert got.dtype == ref.dtype and got.shape == ref.shape
    assert np.array_equal(got, ref), f"Clenshaw kernel not bit-identical: max diff {np.abs(got - ref).max():.3e} (tol {tol:.3e})"
    if len(dims) >= 2 and dims[0] >= 5 and np.prod(dims[1:]) >= 8:   # Float32 runs the same FP64 tensor-core kernels
        poly.set_algo(fc.ALGO_DMMA, dt)
        assert poly.get_algo(False, dt) == fc.ALGO_DMMA
        got2 = poly(pts)
        assert np.abs(got2 - ref).max() <= tol, f"DMMA: {np.abs(got2 - ref).max():.3e} > {tol:.3e}"
        poly.set_algo(fc.ALGO_DMMA_RING, dt)      # 2-D: the streaming kernel instead of the resident one
        got3 = poly(pts)
        assert np.abs(got3 - ref).max() <= tol, f"DMMA ring: {np.abs(got3 - ref).max():.3e} > {tol:.3e}"
    # single-point call == the batched entry (runtests.jl:176-178)
    one = poly(pts[5] if len(dims) > 1 else pts[5, 0])
    poly.set_algo(fc.ALGO_AUTO, dt)
    auto = poly(pts)
    assert np.abs(auto - ref).max() <= tol
    assert np.abs(np.asarray(one) - ref[5]).max() <= tol


@pytest.mark.parametrize("case", CASES, ids=ids)
def test_jacobian_parity(fc, orc, case):
    dims, K, cplx, dt = case
    rng, c, lb, ub = make(*case, seed=12)
    pts = points(rng, lb, ub, 1501, dt)
    v0, J0 = orc.OracleChebPoly(c, lb, ub).jacobian(pts)
    poly = fc.ChebPoly(c, lb, ub)
    poly.set_algo(fc.ALGO_CLENSHAW, dt)
    v, J = fc.chebjacobian(poly, pts)
    assert J.shape == J0.shape == (len(pts), K or 1, len(dims))
    assert np.array_equal(v, v0) and np.array_equal(J, J0), "Clenshaw Jacobian kernel not bit-identical to the oracle"
    for d, n in enumerate(dims):
        if n == 1:
            assert np.all(J[:, :, d] == 0)       # exactly zero (runtests.jl:152,158,165)
    if len(dims) >= 2 and dims[0] >= 5 and np.prod(dims[1:]) >= 8:
        poly.set_algo(fc.ALGO_DMMA, dt)
        v2, J2 = fc.chebjacobian(poly, pts)
        tol = 16 * np.finfo(dt).eps * np.abs(c).sum()
        assert np.abs(v2 - v0).max() <= tol
        for d, n in enumerate(dims):
            told = jac_tol(c, len(dims), d, lb, ub, dt)
            assert np.abs(J2[:, :, d] - J0[:, :, d]).max() <= told, (d, np.abs(J2[:, :, d] - J0[:, :, d]).max(), told)
        poly.set_algo(fc.ALGO_DMMA_RING, dt)
        v3, J3 = fc.chebjacobian(poly, pts)
        assert np.abs(v3 - v0).max() <= tol
        for d, n in enumerate(dims):
            assert np.abs(J3[:, :, d] - J0[:, :, d]).max() <= jac_tol(c, len(dims), d, lb, ub, dt)


@pytest.mark.parametrize("K,npts", [(None, 150_011), (3, 60_007)], ids=["scalar", "K3"])
def test_resident_2d_full_ctas(fc, orc, K, npts):
    """configs[1] shape (order (64,64)) at a batch large enough that the resident 2-D kernel runs with all of its
    warps, so the A-fragment scratch tables are shared under their locks; out-of-domain index from a late unit."""
    case = ((65, 65), K, False, np.float64)
    rng, c, lb, ub = make(*case, seed=21)
    pts = points(rng, lb, ub, npts, np.float64)
    ref = orc.OracleChebPoly(c, lb, ub)
    poly = fc.ChebPoly(c, lb, ub)
    poly.set_algo(fc.ALGO_DMMA)
    tol = 16 * np.finfo(float).eps * np.abs(c).sum()
    v0, J0 = ref.jacobian(pts)
    assert np.abs(poly(pts) - v0).max() <= tol
    v, J = fc.chebjacobian(poly, pts)
    assert np.abs(v - v0).max() <= tol
    for d in range(2):
        assert np.abs(J[:, :, d] - J0[:, :, d]).max() <= jac_tol(c, 2, d, lb, ub)
    bad = pts.copy()
    bad[npts - 3, 1] = ub[1] + 1.0
    bad[npts - 2, 0] = np.nan
    with pytest.raises(fc.ArgumentError) as ei:
        poly(bad)
    assert str(bad[npts - 3]) in str(ei.value)    # the FIRST offending point is reported
    with pytest.raises(fc.ArgumentError):
        fc.chebjacobian(poly, bad)


def test_gradient_shapes_and_errors(fc, orc):
    rng = np.random.default_rng(3)
    c = rng.uniform(-1, 1, (9, 7))
    lb, ub = np.array([0.0, 1.0]), np.array([2.0, 4.0])
    p = fc.ChebPoly(c, lb, ub)
    ref = orc.OracleChebPoly(c, lb, ub)
    x = np.array([0.3, 2.5])
    v, g = fc.chebgradient(p, x)
    v0, g0 = ref.gradient(x)
    assert g.shape == (2,) and v == v0 and np.array_equal(g, g0)
    pv = fc.ChebPoly(rng.uniform(-1, 1, (9, 7, 2)), lb, ub)
    with pytest.raises(fc.DimensionMismatch):
        fc.chebgradient(pv, x)                    # eval.jl:170
    p1 = fc.ChebPoly(rng.uniform(-1, 1, 12), [0.0], [1.0])
    v, dv = fc.chebgradient(p1, 0.25)             # 1-D scalar form returns a scalar derivative (eval.jl:174-177)
    assert np.ndim(v) == 0 and np.ndim(dv) == 0


def test_domain_errors(fc):
    rng = np.random.default_rng(4)
    p = fc.ChebPoly(rng.uniform(-1, 1, (8, 8)), [0.0, 0.0], [1.0, 1.0])
    pts = rng.random((1000, 2))
    p(pts)
    bad = pts.copy()
    bad[637, 1] = 1.0000001
    bad[900, 0] = -0.1
    with pytest.raises(fc.ArgumentError) as ei:
        p(bad)                                    # eval.jl:64; first offending point is reported
    assert "1.0000001" in str(ei.value)
    nan = pts.copy()
    nan[10, 0] = np.nan
    with pytest.raises(fc.ArgumentError):
        p(nan)                                    # NaN fails the closed-box test
    with pytest.raises(fc.ArgumentError):
        fc.chebjacobian(p, bad)                   # eval.jl:148
    with pytest.raises(fc.ArgumentError):
        p(np.array([0.5, 2.0]))
    assert p(np.empty((0, 2))).shape == (0,)      # empty batch


def test_promotion_rules(fc, orc):
    """Float32 polynomial at Float32 points stays Float32; at Float64 points it promotes (eval.jl:25,65)."""
    rng = np.random.default_rng(6)
    c32 = rng.uniform(-1, 1, (10, 6)).astype(np.float32)
    lb, ub = np.array([0, 0], dtype=np.float32), np.array([1, 2], dtype=np.float32)
    p = fc.ChebPoly(c32, lb, ub)
    x32 = (rng.random((50, 2)) * [1, 2]).astype(np.float32)
    assert p(x32).dtype == np.float32
    r = p(x32.astype(np.float64))
    assert r.dtype == np.float64
    assert np.array_equal(r, orc.OracleChebPoly(c32, lb, ub)(x32.astype(np.float64)))
    pint = fc.ChebPoly(c32.astype(np.float64), np.array([0, 0]), np.array([1, 2]))   # Td = Int (runtests.jl:170)
    assert np.array_equal(pint(x32.astype(np.float64)), orc.OracleChebPoly(c32.astype(np.float64), np.array([0, 0]), np.array([1, 2]))(x32.astype(np.float64)))


def test_device_resident_batches(fc, orc):
    """torch CUDA tensors in, torch CUDA tensors out, on torch's current stream (no host round trip)."""
    import torch

    rng = np.random.default_rng(7)
    c = rng.uniform(-1, 1, (12, 11, 10, 3))
    lb, ub = -np.ones(3), np.ones(3)
    pts = rng.uniform(-1, 1, (100_003, 3))
    p = fc.ChebPoly(c, lb, ub)
    ref = orc.OracleChebPoly(c, lb, ub)(pts)
    tol = 16 * np.finfo(float).eps * np.abs(c).sum()
    t = torch.from_numpy(pts).cuda()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        out = p(t)
        v, J = fc.chebjacobian(p, t)
    s.synchronize()
    assert out.is_cuda and out.shape == (len(pts), 3)
    assert np.abs(out.cpu().numpy() - ref).max() <= tol
    assert np.abs(v.cpu().numpy() - ref).max() <= tol
    assert J.shape == (len(pts), 3, 3)
    bad = t.clone()
    bad[77777, 2] = 1.5
    with pytest.raises(fc.ArgumentError):
        p(bad)


def test_tuple_layout(fc, orc):
    """fastcheb_eval_jac_tuple writes Julia's Vector{Tuple{SVector{K},SMatrix{K,N}}} records."""
    import ctypes

    rng = np.random.default_rng(8)
    c = rng.uniform(-1, 1, (7, 6, 2))
    lb, ub = np.zeros(2), np.ones(2)
    pts = rng.random((500, 2))
    p = fc.ChebPoly(c, lb, ub)
    p.set_algo(fc.ALGO_CLENSHAW)
    v0, J0 = orc.OracleChebPoly(c, lb, ub).jacobian(pts)
    rec = np.empty((500, 2 + 4))
    rc = fc._capi.lib().fastcheb_eval_jac_tuple(p._handle(np.float64), pts.ctypes.data, 500, rec.ctypes.data, fc._capi.MEM_HOST, None)
    assert rc == 0
    assert np.array_equal(rec[:, :2], v0)
    assert np.array_equal(rec[:, 2:].reshape(500, 2, 2), np.swapaxes(J0, 1, 2))   # column-major K x N


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_eval_many_bit_identical_to_oracle(fc, orc, dt):
    """configs[4] evaluation leg: many independent 1-D polynomials at their own points in one launch
    (fastcheb_eval_many) == a loop of oracle evaluations, bit for bit (values and chebgradient derivative)."""
    rng = np.random.default_rng(77)
    for n, K, cplx, M in ((65, None, False, 1000), (1, None, False, 7), (2, None, False, 33), (17, 3, False, 130),
                          (40, 2, True, 257), (9, None, True, 64)):
        howmany = 37
        shape = (howmany, n) + ((K,) if K else ())
        c = rng.uniform(-1, 1, shape)
        if cplx:
            c = (c + 1j * rng.uniform(-1, 1, shape)).astype(np.complex128 if dt == np.float64 else np.complex64)
        else:
            c = c.astype(dt)
        lb = rng.uniform(-2, 0, howmany).astype(dt)
        ub = (lb + rng.uniform(0.5, 3, howmany)).astype(dt)
        pts = (lb[:, None] + (ub - lb)[:, None] * rng.random((howmany, M))).astype(dt)
        pts = np.clip(pts, lb[:, None], ub[:, None])
        pts[:, 0], pts[:, -1] = lb, ub                                    # closed box
        v, d = fc.eval_many(c, lb, ub, pts, deriv=True)
        v_only = fc.eval_many(c, lb, ub, pts)
        assert np.array_equal(v, v_only)
        for f in (0, 5, howmany - 1):
            ref = orc.OracleChebPoly(c[f], np.array([lb[f]]), np.array([ub[f]]))
            rv, rJ = ref.jacobian(pts[f].reshape(-1, 1))
            assert np.array_equal(v[f], rv.reshape(v[f].shape)), (n, K, cplx, f)
            assert np.array_equal(d[f], rJ.reshape(d[f].shape)), (n, K, cplx, f)
    # shared bounds + domain error with the flat index of the offending point
    c = rng.uniform(-1, 1, (5, 9)).astype(dt)
    pts = rng.random((5, 11)).astype(dt)
    assert fc.eval_many(c, 0.0, 1.0, pts).shape == (5, 11)
    pts[3, 4] = 1.5
    with pytest.raises(fc.ArgumentError) as ei:
        fc.eval_many(c, 0.0, 1.0, pts)
    assert ei.value.index == 3 * 11 + 4


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_eval_many_large_batches_bit_identical(fc, orc, dt):
    """Batches of hundreds to thousands of scalar polynomials (more polynomials than resident warps, ragged point counts,
    per-polynomial and shared bounds): bit-identical to the oracle, and a sub-batch evaluated alone gives the same bits."""
    rng = np.random.default_rng(79)
    for n, howmany, M, per_poly in ((65, 301, 1000, True), (64, 1200, 130, False), (33, 700, 500, True), (17, 5000, 17, True),
                                    (40, 333, 999, False), (5, 2000, 33, True), (1, 400, 400, True), (2, 400, 400, False)):
        c = rng.uniform(-1, 1, (howmany, n)).astype(dt)
        if per_poly:
            lb = rng.uniform(-2, 0, howmany).astype(dt)
            ub = (lb + rng.uniform(0.5, 3, howmany)).astype(dt)
            lbm, ubm = lb[:, None], ub[:, None]
        else:
            lb, ub = dt(-1.25), dt(2.5)
            lbm, ubm = lb, ub
        pts = np.clip((lbm + (ubm - lbm) * rng.random((howmany, M))).astype(dt), lbm, ubm)
        pts[:, 0], pts[:, -1] = lb, ub                                    # closed box
        v, d = fc.eval_many(c, lb, ub, pts, deriv=True)
        assert np.array_equal(fc.eval_many(c, lb, ub, pts), v)
        # the first 40 polynomials alone: same bits
        sl = slice(0, 40)
        v40, d40 = fc.eval_many(c[sl], lb[sl] if per_poly else lb, ub[sl] if per_poly else ub, pts[sl], deriv=True)
        assert np.array_equal(v[sl], v40) and np.array_equal(d[sl], d40), (n, howmany, M)
        for f in (0, 31, 32, howmany // 2, howmany - 1):
            lf, uf = (lb[f], ub[f]) if per_poly else (lb, ub)
            ref = orc.OracleChebPoly(c[f], np.array([lf]), np.array([uf]))
            rv, rJ = ref.jacobian(pts[f].reshape(-1, 1))
            assert np.array_equal(v[f], rv.reshape(v[f].shape)), (n, howmany, M, f)
            assert np.array_equal(d[f], rJ.reshape(d[f].shape)), (n, howmany, M, f)
    # domain error: the flat index of the first offending point
    c = rng.uniform(-1, 1, (600, 20)).astype(dt)
    pts = rng.random((600, 100)).astype(dt)
    pts[417, 63] = 1.5
    pts[599, 99] = np.nan
    with pytest.raises(fc.ArgumentError) as ei:
        fc.eval_many(c, 0.0, 1.0, pts)
    assert ei.value.index == 417 * 100 + 63


def test_eval_many_device_mode_matches_host_mode(fc):
    """fastcheb_eval_many / fastcheb_eval_many_shared with mem = DEVICE (every pointer, the bounds included, in device
    memory; the call only enqueues; status_dev carries the first out-of-domain index): the same bits as the synchronous
    host-mode call, for per-polynomial and for shared points, values and derivatives."""
    import ctypes
    import torch
    from fastcheb import _capi
    L = _capi.lib()
    rng = np.random.default_rng(81)
    dev = torch.device("cuda", 0)
    for dt, tdt, code in ((np.float64, torch.float64, _capi.F64), (np.float32, torch.float32, _capi.F32)):
        for n, howmany, M in ((65, 300, 1000), (33, 130, 77), (8, 64, 200)):
            c = (rng.uniform(-1, 1, (howmany, n)) * 0.9 ** np.arange(n)).astype(dt)
            lb, ub = -1.5, 0.75
            x_sh = np.clip((lb + (ub - lb) * rng.random(M)).astype(dt), dt(lb), dt(ub))
            x_pp = np.clip((lb + (ub - lb) * rng.random((howmany, M))).astype(dt), dt(lb), dt(ub))
            d_c = torch.from_numpy(c).to(dev)
            d_lb = torch.tensor([lb], dtype=torch.float64, device=dev)
            d_ub = torch.tensor([ub], dtype=torch.float64, device=dev)
            sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for shared, x in ((True, x_sh), (False, x_pp)):
                hv, hd = fc.eval_many(c, dt(lb), dt(ub), x, deriv=True)
                d_x = torch.from_numpy(x).to(dev)
                ov = torch.empty((howmany, M), dtype=tdt, device=dev)
                od = torch.empty((howmany, M), dtype=tdt, device=dev)
                st = torch.zeros(1, dtype=torch.int64, device=dev)
                if shared:
                    rc = L.fastcheb_eval_many_shared(howmany, n, code, 0, 1, d_c.data_ptr(), d_lb.data_ptr(), d_ub.data_ptr(),
                                                     d_x.data_ptr(), M, ov.data_ptr(), od.data_ptr(), st.data_ptr(), _capi.MEM_DEVICE, 0, sp)
                else:
                    rc = L.fastcheb_eval_many(howmany, n, code, 0, 1, d_c.data_ptr(), d_lb.data_ptr(), d_ub.data_ptr(), 0,
                                              d_x.data_ptr(), M, ov.data_ptr(), od.data_ptr(), st.data_ptr(), _capi.MEM_DEVICE, 0, sp)
                assert rc == 0, _capi.last_error()
                torch.cuda.synchronize()
                assert int(st.item()) == 0x7fffffffffffffff
                assert np.array_equal(ov.cpu().numpy(), hv), (dt, n, shared)
                assert np.array_equal(od.cpu().numpy(), hd), (dt, n, shared)
            # an out-of-domain shared point is reported through status_dev
            bad = x_sh.copy()
            bad[M // 2] = dt(2.0)
            d_x = torch.from_numpy(bad).to(dev)
            st = torch.zeros(1, dtype=torch.int64, device=dev)
            rc = L.fastcheb_eval_many_shared(howmany, n, code, 0, 1, d_c.data_ptr(), d_lb.data_ptr(), d_ub.data_ptr(), d_x.data_ptr(), M,
                                             ov.data_ptr(), None, st.data_ptr(), _capi.MEM_DEVICE, 0, sp)
            assert rc == 0
            torch.cuda.synchronize()
            assert int(st.item()) == M // 2


def test_cfg5_fit_then_eval_many(fc):
    """configs[4] end to end at full size: 10^5 order-64 fits (tensor-core DCT-I), then every polynomial evaluated at
    10^3 of its own points (10^8 evaluations) against the analytic functions f_i(x) = sin(a_i x + b_i)."""
    rng = np.random.default_rng(8)
    Mfits, n, M = 100_000, 65, 1000
    a, b = rng.uniform(0.5, 3.0, Mfits), rng.uniform(-1, 1, Mfits)
    grid = np.cos(np.pi * np.arange(n) / (n - 1))
    C = fc.chebcoefs(np.sin(a[:, None] * grid[None, :] + b[:, None]), ndim=1, howmany=Mfits)
    x = rng.uniform(-1, 1, (Mfits, M))
    v, d = fc.eval_many(C, -1.0, 1.0, x, deriv=True)
    assert np.abs(v - np.sin(a[:, None] * x + b[:, None])).max() <= 1e-13
    assert np.abs(d - a[:, None] * np.cos(a[:, None] * x + b[:, None])).max() <= 1e-10


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
def test_eval_many_shared_points_matches_oracle(fc, orc, dt):
    """fastcheb_eval_many_shared: every polynomial at the same points = a matrix product on the FP64 tensor cores.
    Values within 16 eps sum|c| of the oracle's per-polynomial Clenshaw evaluation, derivatives within the
    differentiated-series bound; ragged row blocks / point tiles, vector and complex elements, n = 1, 2, 65."""
    rng = np.random.default_rng(78)
    eps = np.finfo(dt).eps
    for n, K, cplx, howmany, M in ((65, None, False, 301, 1000), (65, 3, False, 50, 130), (33, None, True, 77, 257), (1, None, False, 9, 5),
                                   (2, 2, False, 20, 64), (40, 2, True, 130, 63),
                                   # k-step counts between the instantiated ones (n = 19: 5 of 8, n = 37: 9 of 12, n = 51: 13 of 16)
                                   (5, None, False, 3, 1), (19, 3, True, 758, 949), (37, None, False, 100, 64), (51, 2, False, 70, 200)):
        shape = (howmany, n) + ((K,) if K else ())
        # coefficients that decay like those of fitted smooth functions (what configs[4] evaluates).  For dense-random
        # coefficient vectors the Clenshaw recurrence itself carries ~k eps |c_k| of rounding at the end points, and a
        # direct sum (exact there) then differs from it by up to ~1.2 tolerances.
        decay = (0.85 ** np.arange(n)).reshape((1, n) + ((1,) if K else ()))
        c = rng.uniform(-1, 1, shape) * decay
        if cplx:
            c = (c + 1j * rng.uniform(-1, 1, shape) * decay).astype(np.complex128 if dt == np.float64 else np.complex64)
        else:
            c = c.astype(dt)
        lb, ub = dt(-1.5), dt(0.75)
        x = (lb + (ub - lb) * rng.random(M)).astype(dt)
        x = np.clip(x, lb, ub)
        x[0], x[-1] = lb, ub
        v, d = fc.eval_many(c, lb, ub, x, deriv=True)
        assert v.shape == (howmany, M) + ((K,) if K else ()) and v.dtype == c.dtype
        assert np.array_equal(fc.eval_many(c, lb, ub, x), v)
        k2 = np.arange(n, dtype=np.float64) ** 2
        for f in (0, howmany // 2, howmany - 1):
            ref = orc.OracleChebPoly(c[f], np.array([lb]), np.array([ub]))
            rv, rJ = ref.jacobian(x.reshape(-1, 1))
            tol = 16 * eps * np.abs(c[f]).sum()
            assert np.abs(v[f] - rv.reshape(v[f].shape)).max() <= tol, (n, K, cplx, f)
            told = 16 * eps * (np.abs(c[f]).reshape(n, -1) * k2[:, None]).sum() * 2 / (float(ub) - float(lb))
            assert np.abs(d[f] - rJ.reshape(d[f].shape)).max() <= max(told, 0.0), (n, K, cplx, f)
    with pytest.raises(fc.ArgumentError) as ei:
        bad = x.copy()
        bad[17] = 2.0
        fc.eval_many(c, lb, ub, bad)
    assert ei.value.index == 17

"""Memory safety of the vector-store kernels (compute-sanitizer is closed on the GPU pool): results are written into the
middle of sentinel-filled device buffers through the raw C ABI; every batch size around the tile and vector boundaries must
leave the guard bands before and after the result untouched and fill exactly npts * E (and npts * N * E) elements.
Covers the low-order 1-D / N-d kernels, the product-basis kernel, the generic Clenshaw kernel and eval_many."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GUARD = 64  # elements on either side (a multiple of 16 bytes in both precisions)


def _eval_guarded(fc, torch, p, dt, x_np, jac, off_elems=0):
    from fastcheb import _capi
    L = _capi.lib()
    tdt = torch.float64 if dt == np.float64 else torch.float32
    N, E = p.ndim, (p.K or 1) * (2 if p.is_complex else 1)
    npts = x_np.shape[0]
    xin = torch.zeros(GUARD + off_elems + npts * N + GUARD, dtype=tdt, device="cuda")
    xin[GUARD + off_elems:GUARD + off_elems + npts * N] = torch.from_numpy(np.ascontiguousarray(x_np).reshape(-1)).cuda()
    sent = 123456.75
    ov = torch.full((2 * GUARD + off_elems + npts * E,), sent, dtype=tdt, device="cuda")
    oj = torch.full((2 * GUARD + off_elems + npts * N * E,), sent, dtype=tdt, device="cuda")
    status = torch.zeros(1, dtype=torch.int64, device="cuda")
    es = ov.element_size()
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    h = p._handle(dt)
    base = (GUARD + off_elems) * es
    if jac:
        rc = L.fastcheb_eval_jac_async(h, xin.data_ptr() + base, npts, ov.data_ptr() + base, oj.data_ptr() + base, status.data_ptr(), sp)
    else:
        rc = L.fastcheb_eval_async(h, xin.data_ptr() + base, npts, ov.data_ptr() + base, status.data_ptr(), sp)
    assert rc == 0, _capi.last_error()
    torch.cuda.synchronize()
    assert int(status.item()) == 0x7FFFFFFFFFFFFFFF
    ovn, ojn = ov.cpu().numpy(), oj.cpu().numpy()
    lo = GUARD + off_elems
    assert np.all(ovn[:lo] == sent) and np.all(ovn[lo + npts * E:] == sent), "value guard band overwritten"
    assert not np.any(ovn[lo:lo + npts * E] == sent), "value result not fully written"
    if jac:
        assert np.all(ojn[:lo] == sent) and np.all(ojn[lo + npts * N * E:] == sent), "Jacobian guard band overwritten"
        assert not np.any(ojn[lo:lo + npts * N * E] == sent), "Jacobian result not fully written"
    return ovn[lo:lo + npts * E], ojn[lo:lo + npts * N * E]


CASES = [
    ("lo1d", (5,), None), ("lo1d-K3", (7,), 3), ("lo1d-long", (65,), None), ("lo1d-513", (513,), None),
    ("lond-2d", (5, 5), None), ("lond-3d-K2", (3, 4, 3), 2), ("generic-2d", (40, 9), None), ("generic-1d", (600,), None),
]


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("name,shape,K", CASES, ids=[c[0] for c in CASES])
def test_guard_bands(fc, orc, name, shape, K, dt):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(sum(name.encode()))
    N = len(shape)
    c = rng.uniform(-1, 1, shape + ((K,) if K else ())).astype(dt)
    lb, ub = -np.ones(N, dtype=dt), np.ones(N, dtype=dt)
    p = fc.ChebPoly(c, lb, ub)
    o = orc.OracleChebPoly(c, lb, ub)
    for npts in (4096, 4097, 4099, 5119, 5120, 5121, 8191, 8193):
        x = np.clip((2 * rng.random((npts, N)) - 1).astype(dt), -1, 1)
        xe = x if N > 1 else x[:, 0]
        v0, J0 = o.jacobian(xe)
        for jac in (False, True):
            for off in (0, 1):   # off = 1: the buffers are not 16-byte aligned (the scalar paths / the generic kernel)
                v, J = _eval_guarded(fc, torch, p, dt, x, jac, off)
                assert np.array_equal(v.reshape(np.asarray(v0).shape), v0), (name, npts, jac, off)
                if jac:
                    Jr = J.reshape(npts, N, -1).transpose(0, 2, 1).reshape(np.asarray(J0).shape)
                    assert np.array_equal(Jr, J0), (name, npts, off)


def test_guard_bands_product_basis(fc, orc):
    torch = pytest.importorskip("torch")
    xg = orc.chebpoints(64, -1.0, 1.0)
    for dt in (np.float64, np.float32):
        for K in (None, 3):
            f = np.sin(2.3 * xg + 0.4) if K is None else np.stack([np.sin((2.0 + k) * xg + 0.1 * k) for k in range(K)], axis=-1)
            c = orc.chebcoefs(f, 1).astype(dt)
            p = fc.ChebPoly(c, np.array([-1.0], dtype=dt), np.array([1.0], dtype=dt))
            rng = np.random.default_rng(3)
            for npts in (8192, 8193, 8195, 9999):
                x = np.clip((2 * rng.random((npts, 1)) - 1).astype(dt), -1, 1)
                for jac in (False, True):
                    for off in (0, 1):
                        _eval_guarded(fc, torch, p, dt, x, jac, off)
            assert p.get_algo(False, dt) == fc.ALGO_PRODUCT


def test_guard_bands_eval_many(fc, orc):
    torch = pytest.importorskip("torch")
    from fastcheb import _capi
    L = _capi.lib()
    rng = np.random.default_rng(5)
    sent = 98765.5
    for dt, tdt, code in ((np.float64, torch.float64, _capi.F64), (np.float32, torch.float32, _capi.F32)):
        for howmany, n, M in ((37, 33, 1000), (37, 33, 1001), (19, 65, 257), (8, 9, 4)):
            coefs = torch.from_numpy(rng.uniform(-1, 1, (howmany, n)).astype(dt)).cuda()
            x = torch.from_numpy(np.clip((2 * rng.random((howmany, M)) - 1).astype(dt), -1, 1)).cuda()
            lb = torch.full((1,), -1.0, dtype=torch.float64, device="cuda")
            ub = torch.full((1,), 1.0, dtype=torch.float64, device="cuda")
            ov = torch.full((2 * GUARD + howmany * M,), sent, dtype=tdt, device="cuda")
            od = torch.full((2 * GUARD + howmany * M,), sent, dtype=tdt, device="cuda")
            st = torch.zeros(1, dtype=torch.int64, device="cuda")
            sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            base = GUARD * ov.element_size()
            rc = L.fastcheb_eval_many(howmany, n, code, 0, 1, coefs.data_ptr(), lb.data_ptr(), ub.data_ptr(), 0, x.data_ptr(), M,
                                      ov.data_ptr() + base, od.data_ptr() + base, st.data_ptr(), _capi.MEM_DEVICE, 0, sp)
            assert rc == 0, _capi.last_error()
            torch.cuda.synchronize()
            for buf in (ov.cpu().numpy(), od.cpu().numpy()):
                assert np.all(buf[:GUARD] == sent) and np.all(buf[GUARD + howmany * M:] == sent)
                assert not np.any(buf[GUARD:GUARD + howmany * M] == sent)
            cn, xn = coefs.cpu().numpy(), x.cpu().numpy()
            vals = ov.cpu().numpy()[GUARD:GUARD + howmany * M].reshape(howmany, M)
            for i in (0, howmany - 1):
                ref = orc.OracleChebPoly(cn[i], np.array([-1.0], dtype=dt), np.array([1.0], dtype=dt))(xn[i])
                assert np.array_equal(vals[i], ref)

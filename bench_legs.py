"""The `legs` object of bench.py's JSON line: short device-timed runs of every other BASELINE.json config on one GPU,
each with its own roofline, so that the driver-run record carries the 2-D, 1-D, 4-D, Jacobian and fit numbers too.

Every leg: inputs resident in HBM, >= 3 warm-up launches, CUDA events on the launching stream, working sets larger
than the 126 MB L2 (or rotating buffer sets for the small fit launches).  Evaluation legs report algorithmic FP64
flops (W_val = 2 K cplx S, W_jac = 2 K cplx sum_d (d+1) S_d) against the measured FP64 peak
(profiles/r01_micro_fp64.jsonl) and the bytes of points in + results out against the measured HBM copy bandwidth
(MEASURED_PEAKS.json); the `bound` named is the one the shape is limited by.  Fit legs report 2 sizeof(T) prod(n)
bytes per fit against the HBM figure."""
from __future__ import annotations

import ctypes
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _peaks():
    import bench

    fp64, _ = bench.fp64_peak_tflops()
    hbm, _ = bench.hbm_peak_gbs()
    return fp64, hbm


def _work(dims, E, jac):
    N = len(dims)
    Sd = [int(np.prod(dims[d:])) for d in range(N)]
    return 2 * E * (sum((d + 2) * Sd[d] for d in range(N)) if jac else sum(Sd))


def _time(torch, fn, reps, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def eval_leg(torch, fastcheb, _capi, device, dims, K, cplx, dt, npts, jac, reps, coefs=None, lb=None, ub=None):
    L = _capi.lib()
    dev = torch.device("cuda", device)
    fp64, hbm = _peaks()
    rng = np.random.default_rng(1)
    N = len(dims)
    if coefs is None:
        shape = tuple(dims) + ((K,) if K else ())
        c = rng.uniform(-1, 1, shape)
        if cplx:
            c = (c + 1j * rng.uniform(-1, 1, shape)).astype(np.complex128 if dt == np.float64 else np.complex64)
        else:
            c = c.astype(dt)
    else:
        c = coefs
    lb = -np.ones(N, dtype=dt) if lb is None else np.asarray(lb, dtype=dt)
    ub = np.ones(N, dtype=dt) if ub is None else np.asarray(ub, dtype=dt)
    E = (K or 1) * (2 if cplx else 1)
    poly = fastcheb.ChebPoly(c, lb, ub, device=device)
    tdt = torch.float64 if dt == np.float64 else torch.float32
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    pts = torch.rand((npts, N), dtype=tdt, device=dev, generator=g) * torch.tensor(ub - lb, dtype=tdt, device=dev) + torch.tensor(lb, dtype=tdt, device=dev)
    pts = torch.minimum(torch.maximum(pts, torch.tensor(lb, dtype=tdt, device=dev)), torch.tensor(ub, dtype=tdt, device=dev))
    out_v = torch.empty((npts, E), dtype=tdt, device=dev)
    out_J = torch.empty((npts, N, E), dtype=tdt, device=dev) if jac else None
    status = torch.zeros(1, dtype=torch.int64, device=dev)
    h = poly._handle(dt)
    sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def run(_):
        if jac:
            rc = L.fastcheb_eval_jac_async(h, pts.data_ptr(), npts, out_v.data_ptr(), out_J.data_ptr(), status.data_ptr(), sp)
        else:
            rc = L.fastcheb_eval_async(h, pts.data_ptr(), npts, out_v.data_ptr(), status.data_ptr(), sp)
        if rc != 0:
            raise RuntimeError(_capi.last_error())
    ms = _time(torch, run, reps)
    if int(status.item()) != 0x7FFFFFFFFFFFFFFF:
        raise RuntimeError("out-of-domain point in a synthetic leg batch")
    rate = npts / (ms * 1e-3)
    W = _work(dims, E, jac)
    rs = np.dtype(dt).itemsize
    bytes_pp = N * rs + E * rs * (1 + (N if jac else 0))
    algo = {1: "clenshaw", 2: "dmma", 4: "product-basis"}.get(poly.get_algo(jac, dt), "?")
    # Float32 polynomials on the tensor-core path compute in FP64; the CUDA-core Float32 kernels have twice the peak
    fpeak = fp64 if (dt == np.float64 or algo == "dmma") else 2 * fp64
    tf, gbs = rate * W / 1e12, rate * bytes_pp / 1e9
    bound = "tensor" if tf / fpeak >= gbs / hbm else "hbm"
    roof = ({"bound": "tensor", "achieved": tf, "peak": fpeak, "unit": "TFLOP/s", "frac": tf / fpeak} if bound == "tensor"
            else {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm})
    roof.update({"flops_per_point": W, "bytes_per_point": bytes_pp, "frac_fp": tf / fpeak, "frac_hbm": gbs / hbm})
    poly.close()
    del pts, out_v, out_J
    return {"value": rate, "unit": "points/s", "ms": ms, "points": npts, "algo": algo, "dtype": np.dtype(dt).name, "roofline": roof}


def eval_many_leg(torch, _capi, device, dt, der, reps, shared=False, howmany=100_000, n=65, M=1000):
    L = _capi.lib()
    dev = torch.device("cuda", device)
    fp64, hbm = _peaks()
    tdt, code = (torch.float64, _capi.F64) if dt == np.float64 else (torch.float32, _capi.F32)
    g = torch.Generator(device=dev)
    g.manual_seed(9)
    coefs = torch.rand((howmany, n), dtype=tdt, device=dev, generator=g) * 2 - 1
    pts = torch.rand((1 if shared else howmany, M), dtype=tdt, device=dev, generator=g) * 2 - 1
    lb = torch.full((1,), -1.0, dtype=torch.float64, device=dev)
    ub = torch.full((1,), 1.0, dtype=torch.float64, device=dev)
    out = torch.empty((howmany, M), dtype=tdt, device=dev)
    outd = torch.empty((howmany, M), dtype=tdt, device=dev) if der else None
    status = torch.zeros(1, dtype=torch.int64, device=dev)
    sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def run(_):
        if shared:
            rc = L.fastcheb_eval_many_shared(howmany, n, code, 0, 1, coefs.data_ptr(), lb.data_ptr(), ub.data_ptr(), pts.data_ptr(), M,
                                             out.data_ptr(), outd.data_ptr() if der else None, status.data_ptr(), _capi.MEM_DEVICE,
                                             device, sp)
        else:
            rc = L.fastcheb_eval_many(howmany, n, code, 0, 1, coefs.data_ptr(), lb.data_ptr(), ub.data_ptr(), 0, pts.data_ptr(), M,
                                      out.data_ptr(), outd.data_ptr() if der else None, status.data_ptr(), _capi.MEM_DEVICE, device, sp)
        if rc != 0:
            raise RuntimeError(_capi.last_error())
    ms = _time(torch, run, reps)
    rate = howmany * M / (ms * 1e-3)
    rs = np.dtype(dt).itemsize
    W = 2 * n * (2 if der else 1)
    bpp = rs * ((1 if shared else 2) + (1 if der else 0))
    fpeak = fp64 if (dt == np.float64 or shared) else 2 * fp64
    tf, gbs = rate * W / 1e12, rate * bpp / 1e9
    roof = {"bound": "tensor", "achieved": tf, "peak": fpeak, "unit": "TFLOP/s", "frac": tf / fpeak, "flops_per_point": W,
            "frac_fp": tf / fpeak, "frac_hbm": gbs / hbm}
    return {"value": rate, "unit": "points/s", "ms": ms, "points": howmany * M, "dtype": np.dtype(dt).name,
            "polynomials": howmany, "points_per_polynomial": M, "shared_points": shared, "roofline": roof}


def fit_leg(torch, _capi, device, dims, howmany, dt, reps):
    L = _capi.lib()
    dev = torch.device("cuda", device)
    _, hbm = _peaks()
    tdt, code = (torch.float64, _capi.F64) if dt == np.float64 else (torch.float32, _capi.F32)
    per = int(np.prod(dims))
    rs = np.dtype(dt).itemsize
    bytes_per_fit = 2 * rs * per
    nbuf = min(8, max(2, int(np.ceil(300e6 / (howmany * bytes_per_fit))) + 1))  # no launch finds its arrays in the L2
    g = torch.Generator(device=dev)
    g.manual_seed(8)
    vals = [torch.rand((howmany, per), dtype=tdt, device=dev, generator=g) * 2 - 1 for _ in range(nbuf)]
    outs = [torch.empty((howmany, per), dtype=tdt, device=dev) for _ in range(nbuf)]
    d64 = _capi.i64_array(dims)
    sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def run(i):
        rc = L.fastcheb_fit_async(len(dims), d64, code, 0, 1, howmany, vals[i % nbuf].data_ptr(), outs[i % nbuf].data_ptr(), None, None,
                                  device, sp)
        if rc != 0:
            raise RuntimeError(_capi.last_error())
    l0 = L.fastcheb_launch_count()
    ms = _time(torch, run, reps, warm=3)
    launches = (L.fastcheb_launch_count() - l0) / (reps + 3)
    gbs = howmany * bytes_per_fit / (ms * 1e-3) / 1e9
    return {"value": howmany / (ms * 1e-3), "unit": "fits/s", "ms": ms, "fits": howmany, "dtype": np.dtype(dt).name,
            "launches_per_call": launches,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "bytes_per_fit": bytes_per_fit}}


def run_all(device: int = 0, quick: bool = False):
    import torch

    import fastcheb
    from fastcheb import _capi

    torch.cuda.set_device(device)
    reps = 2 if quick else 5
    legs = {}

    def guard(name, fn):
        try:
            legs[name] = fn()
        except Exception as e:  # a leg must never take the headline line down with it
            legs[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()

    f64, f32 = np.float64, np.float32
    # configs[0]: 1-D order 200, coefficients = the README fit of sin(2x + 3cos(4x)) on [0, 10] (tol = 0)
    xg = fastcheb.chebpoints(200, 0.0, 10.0)
    c1 = np.asarray(fastcheb.chebinterp(np.sin(2 * xg + 3 * np.cos(4 * xg)), 0.0, 10.0, tol=0.0, device=device).coefs)
    for jac in (False, True):
        sfx = "_grad" if jac else "_val"
        guard("cfg0_eval1d_order200" + sfx, lambda: eval_leg(torch, fastcheb, _capi, device, (201,), None, False, f64, 1 << 24, jac, reps,
                                                               coefs=c1, lb=[0.0], ub=[10.0]))
        guard("cfg1_eval2d_order64" + sfx, lambda: eval_leg(torch, fastcheb, _capi, device, (65, 65), None, False, f64, 1 << 24, jac, reps))
    guard("cfg2_eval3d_scalar_val", lambda: eval_leg(torch, fastcheb, _capi, device, (33, 33, 33), None, False, f64, 1 << 23, False, reps))
    guard("cfg3_eval4d_complex_val", lambda: eval_leg(torch, fastcheb, _capi, device, (17,) * 4, None, True, f64, 1 << 22, False, reps))
    guard("cfg3_eval4d_complex_jac", lambda: eval_leg(torch, fastcheb, _capi, device, (17,) * 4, None, True, f64, 1 << 21, True, reps))
    guard("eval2d_order32_val", lambda: eval_leg(torch, fastcheb, _capi, device, (33, 33), None, False, f64, 1 << 24, False, reps))
    # configs[4] evaluates FITTED order-64 polynomials: coefficients of an analytic f_i(x) = sin(a x + b),
    # which the product-basis path accepts; the dense-random variant (no decay: AUTO keeps the bit-exact Clenshaw
    # kernel, see eval_pb1d.cuh) is reported next to it
    x64 = fastcheb.chebpoints(64, -1.0, 1.0)
    c64 = np.asarray(fastcheb.chebinterp(np.sin(2.3 * x64 + 0.4), -1.0, 1.0, tol=0.0, device=device).coefs)
    guard("cfg4_eval1d_order64_f64_val", lambda: eval_leg(torch, fastcheb, _capi, device, (65,), None, False, f64, 1 << 25, False, reps, coefs=c64))
    guard("cfg4_eval1d_order64_f32_val", lambda: eval_leg(torch, fastcheb, _capi, device, (65,), None, False, f32, 1 << 25, False, reps,
                                                          coefs=c64.astype(np.float32)))
    guard("cfg4_eval1d_order64_dense_random_f64_val", lambda: eval_leg(torch, fastcheb, _capi, device, (65,), None, False, f64, 1 << 25, False, reps))
    for dt, nm in ((f64, "f64"), (f32, "f32")):
        guard(f"cfg4_eval_many_{nm}_val", lambda: eval_many_leg(torch, _capi, device, dt, False, reps))
        if hasattr(_capi.lib(), "fastcheb_eval_many_shared"):
            guard(f"cfg4_eval_many_shared_points_{nm}_val", lambda: eval_many_leg(torch, _capi, device, dt, False, reps, shared=True))
    guard("cfg4_eval_many_f64_grad", lambda: eval_many_leg(torch, _capi, device, f64, True, reps))
    # low-order, memory-bound evaluations (north_star: "achieved HBM GB/s ... for the fit and the low-order, memory-bound
    # evaluations"): a point's coordinates and results cost more than its arithmetic (eval_lo1d.cuh / eval_lond.cuh)
    guard("loworder_eval1d_order4_f64_val", lambda: eval_leg(torch, fastcheb, _capi, device, (5,), None, False, f64, 1 << 26, False, reps))
    guard("loworder_eval1d_order4_f32_val", lambda: eval_leg(torch, fastcheb, _capi, device, (5,), None, False, f32, 1 << 27, False, reps))
    guard("loworder_eval1d_order8_f64_val", lambda: eval_leg(torch, fastcheb, _capi, device, (9,), None, False, f64, 1 << 26, False, reps))
    guard("loworder_eval1d_order8_f64_grad", lambda: eval_leg(torch, fastcheb, _capi, device, (9,), None, False, f64, 1 << 26, True, reps))
    guard("loworder_eval2d_order4_f64_val", lambda: eval_leg(torch, fastcheb, _capi, device, (5, 5), None, False, f64, 1 << 25, False, reps))
    guard("loworder_eval2d_order4_f64_grad", lambda: eval_leg(torch, fastcheb, _capi, device, (5, 5), None, False, f64, 1 << 25, True, reps))
    freps = 10 if quick else 40
    for dt, nm in ((f64, "f64"), (f32, "f32")):
        guard(f"cfg4_fit1d_order64_{nm}_1e5", lambda: fit_leg(torch, _capi, device, (65,), 100_000, dt, freps))
        guard(f"cfg4_fit1d_order64_{nm}_2^20", lambda: fit_leg(torch, _capi, device, (65,), 1 << 20, dt, freps // 2))
    guard("fit2d_65x65_f64_x4096", lambda: fit_leg(torch, _capi, device, (65, 65), 4096, f64, reps))
    guard("fit2d_65x65_f32_x4096", lambda: fit_leg(torch, _capi, device, (65, 65), 4096, f32, reps))
    guard("fit2d_33x33_f64_x65536", lambda: fit_leg(torch, _capi, device, (33, 33), 65536, f64, reps))
    guard("fit2d_33x33_f32_x65536", lambda: fit_leg(torch, _capi, device, (33, 33), 65536, f32, reps))
    guard("fit3d_33x33x33_f64_x512", lambda: fit_leg(torch, _capi, device, (33, 33, 33), 512, f64, reps))
    guard("fit3d_33x33x33_f32_x512", lambda: fit_leg(torch, _capi, device, (33, 33, 33), 512, f32, reps))
    guard("fit1d_n129_f64_x2^18", lambda: fit_leg(torch, _capi, device, (129,), 1 << 18, f64, reps))
    guard("fit1d_n201_f64_x2^17", lambda: fit_leg(torch, _capi, device, (201,), 1 << 17, f64, reps))
    guard("fit1d_n257_f64_x2^17", lambda: fit_leg(torch, _capi, device, (257,), 1 << 17, f64, reps))
    return legs


if __name__ == "__main__":
    print(json.dumps(run_all(0, quick=bool(os.environ.get("QUICK"))), indent=1))

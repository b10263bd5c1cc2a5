#!/usr/bin/env python
"""Evaluation throughput of every BASELINE.json config shape (device-resident points, CUDA events), value and
value+Jacobian, tensor-core and Clenshaw paths — the numbers behind DESIGN.md's per-config table.
    python tools/bench_configs.py > profiles/rNN_configs.jsonl        (one JSON line per measurement)"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastcheb  # noqa: E402
from fastcheb import _capi  # noqa: E402

CONFIGS = [
    ("cfg1: 1-D order 200 real F64", (201,), None, False, np.float64, 1 << 24),
    ("cfg2: 2-D order (64,64) real F64", (65, 65), None, False, np.float64, 1 << 24),
    ("2-D order (64,64) SVector{3} F64", (65, 65), 3, False, np.float64, 1 << 23),
    ("2-D order (32,32) real F64", (33, 33), None, False, np.float64, 1 << 24),
    ("cfg3: 3-D order (32,32,32) SVector{3} F64", (33, 33, 33), 3, False, np.float64, 1 << 23),
    ("cfg3 scalar: 3-D order (32,32,32) real F64", (33, 33, 33), None, False, np.float64, 1 << 23),
    ("cfg4: 4-D order (16,16,16,16) ComplexF64", (17,) * 4, None, True, np.float64, 1 << 22),
    ("cfg5 eval: 1-D order 64 real F64", (65,), None, False, np.float64, 1 << 25),
    ("cfg5 eval: 1-D order 64 real F32", (65,), None, False, np.float32, 1 << 25),
    ("3-D order (32,32,32) SVector{3} F32", (33, 33, 33), 3, False, np.float32, 1 << 22),
]


# extra real-Float64 shapes from the environment (AUTO-threshold checks): BENCH_SHAPES="9x9,17x17x17"
for _s in filter(None, os.environ.get("BENCH_SHAPES", "").split(",")):
    _d = tuple(int(v) for v in _s.split("x"))
    CONFIGS.append((f"shape {_s} real F64", _d, None, False, np.float64, 1 << 22))


def work(dims, E, jac):
    N = len(dims)
    Sd = [int(np.prod(dims[d:])) for d in range(N)]
    return 2 * E * (sum((d + 2) * Sd[d] for d in range(N)) if jac else sum(Sd))


def main():
    L = _capi.lib()
    dev = torch.device("cuda", 0)
    peak64 = 37.0
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6400.0
    only = os.environ.get("BENCH_ONLY")
    for name, dims, K, cplx, dt, npts in CONFIGS:
        if only and only not in name:
            continue
        npts = int(os.environ.get("BENCH_NPTS", npts))   # e.g. the literal 10^6 points of configs[0]
        rng = np.random.default_rng(1)
        shape = dims + ((K,) if K else ())
        c = rng.uniform(-1, 1, shape)
        if cplx:
            c = c + 1j * rng.uniform(-1, 1, shape)
            c = c.astype(np.complex128 if dt == np.float64 else np.complex64)
        else:
            c = c.astype(dt)
        N = len(dims)
        E = (K or 1) * (2 if cplx else 1)
        poly = fastcheb.ChebPoly(c, -np.ones(N, dtype=dt), np.ones(N, dtype=dt))
        tdt = torch.float64 if dt == np.float64 else torch.float32
        pts = torch.rand((npts, N), dtype=tdt, device=dev) * 2 - 1
        out_v = torch.empty((npts, E), dtype=tdt, device=dev)
        out_J = torch.empty((npts, N, E), dtype=tdt, device=dev)
        status = torch.zeros(1, dtype=torch.int64, device=dev)
        h = poly._handle(dt)
        sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for algo in ("dmma", "dmma_ring", "clenshaw"):
            if algo == "dmma_ring" and N != 2:   # only 2-D has a second tensor-core kernel (resident vs ring)
                continue
            try:
                poly.set_algo({"dmma": fastcheb.ALGO_DMMA, "dmma_ring": fastcheb.ALGO_DMMA_RING,
                               "clenshaw": fastcheb.ALGO_CLENSHAW}[algo], rdt=dt)
            except Exception:
                continue
            for jac in (False, True):
                def run():
                    if jac:
                        return L.fastcheb_eval_jac_async(h, pts.data_ptr(), npts, out_v.data_ptr(), out_J.data_ptr(), status.data_ptr(), sp)
                    return L.fastcheb_eval_async(h, pts.data_ptr(), npts, out_v.data_ptr(), status.data_ptr(), sp)
                for _ in range(3):
                    assert run() == 0, _capi.last_error()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 5
                e0.record()
                for _ in range(reps):
                    run()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                rate = npts / (ms * 1e-3)
                W = work(dims, E, jac)
                rs = np.dtype(dt).itemsize
                bytes_pp = N * rs + E * rs * (1 + (N if jac else 0))
                peak = peak64 if (dt == np.float64 or algo == "dmma") else 2 * peak64  # Float32 on the DMMA path computes in FP64
                print(json.dumps({"config": name, "algo": algo, "mode": "jac" if jac else "val", "points": npts, "ms": ms,
                                  "points_per_s": rate, "tflops_algorithmic": rate * W / 1e12,
                                  "frac_of_fp_peak": rate * W / 1e12 / peak, "peak_tflops": peak,
                                  "gb_per_s": rate * bytes_pp / 1e9, "frac_of_hbm": rate * bytes_pp / 1e9 / hbm}), flush=True)
        poly.close()
        del pts, out_v, out_J


def many():
    """configs[4] evaluation leg: 10^5 order-64 polynomials x 10^3 points each (fastcheb_eval_many, device-resident)."""
    L = _capi.lib()
    dev = torch.device("cuda", 0)
    hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6400.0
    howmany, n, M = 100_000, 65, 1000
    sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    for dt, tdt, code, peak in ((np.float64, torch.float64, _capi.F64, 37.0), (np.float32, torch.float32, _capi.F32, 74.0)):
        coefs = torch.rand((howmany, n), dtype=tdt, device=dev) * 2 - 1
        pts = torch.rand((howmany, M), dtype=tdt, device=dev) * 2 - 1
        lb = torch.full((1,), -1.0, dtype=torch.float64, device=dev)
        ub = torch.full((1,), 1.0, dtype=torch.float64, device=dev)
        out = torch.empty((howmany, M), dtype=tdt, device=dev)
        outd = torch.empty((howmany, M), dtype=tdt, device=dev)
        status = torch.zeros(1, dtype=torch.int64, device=dev)
        for der in (False, True):
            def run():
                return L.fastcheb_eval_many(howmany, n, code, 0, 1, coefs.data_ptr(), lb.data_ptr(), ub.data_ptr(), 0,
                                            pts.data_ptr(), M, out.data_ptr(), outd.data_ptr() if der else None,
                                            status.data_ptr(), _capi.MEM_DEVICE, 0, sp)
            for _ in range(3):
                assert run() == 0, _capi.last_error()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            rate = howmany * M / (ms * 1e-3)
            rs = np.dtype(dt).itemsize
            W = 2 * n * (2 if der else 1)  # W_jac = 2 K (d+1) S_d
            bpp = rs * (2 + (1 if der else 0))
            print(json.dumps({"config": f"cfg5 eval_many: 1e5 x order-64 {np.dtype(dt).name}, 1e3 points each", "algo": "clenshaw",
                              "mode": "grad" if der else "val", "points": howmany * M, "ms": ms, "points_per_s": rate,
                              "tflops_algorithmic": rate * W / 1e12, "frac_of_fp_peak": rate * W / 1e12 / peak,
                              "peak_tflops": peak, "gb_per_s": rate * bpp / 1e9, "frac_of_hbm": rate * bpp / 1e9 / hbm}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "many":
        many()
    else:
        main()
        many()

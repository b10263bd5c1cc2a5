#!/usr/bin/env python
"""Throughput of batched fits other than the order-64 lines of configs[4]: batches of 2-D / 3-D arrays and longer
lines, device-resident (fastcheb_fit_async), CUDA events.  One JSON line per shape: fits/s, algorithmic GB/s
(2 sizeof(T) prod(n) per fit) and the fraction of the measured HBM copy bandwidth.
    python tools/probe_fit_nd.py > profiles/rNN_fit_nd.jsonl"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastcheb import _capi  # noqa: E402

SHAPES = [  # dims, howmany, dtype
    ((65, 65), 4096, np.float64),
    ((33, 33), 65536, np.float64),
    ((33, 33), 65536, np.float32),
    ((17, 17, 17), 8192, np.float64),
    ((65, 65), 4096, np.float32),
    ((33, 33, 33), 512, np.float64),
    ((33, 33, 33), 512, np.float32),
    ((17, 17, 17, 17), 128, np.float64),
    ((129,), 1 << 18, np.float64),
    ((201,), 1 << 17, np.float64),
    ((257,), 1 << 17, np.float64),
    ((257,), 1 << 17, np.float32),
    ((33,), 1 << 21, np.float64),
    ((65,), 100_000, np.float64),
    ((65,), 100_000, np.float32),
    ((65,), 1 << 20, np.float64),
    ((65,), 1 << 20, np.float32),
]


def main():
    L = _capi.lib()
    dev = torch.device("cuda", 0)
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm = 6400.0
    sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    only = os.environ.get("FIT_ONLY")
    for dims, howmany, dt in SHAPES:
        tag = "x".join(map(str, dims)) + "-" + np.dtype(dt).name
        if only and only not in tag:
            continue
        if os.environ.get("FIT_HOWMANY"):
            howmany = int(os.environ["FIT_HOWMANY"])
        tdt = torch.float64 if dt == np.float64 else torch.float32
        per = int(np.prod(dims))
        nbuf = max(2, int(np.ceil(300e6 / (howmany * per * 2 * np.dtype(dt).itemsize))) + 1)
        nbuf = min(nbuf, 8)
        vals = [torch.rand((howmany, per), dtype=tdt, device=dev) * 2 - 1 for _ in range(nbuf)]
        outs = [torch.empty((howmany, per), dtype=tdt, device=dev) for _ in range(nbuf)]
        d64 = _capi.i64_array(dims)
        code = _capi.F64 if dt == np.float64 else _capi.F32

        def run(i):
            rc = L.fastcheb_fit_async(len(dims), d64, code, 0, 1, howmany, vals[i % nbuf].data_ptr(), outs[i % nbuf].data_ptr(),
                                      None, None, 0, sp)
            assert rc == 0, _capi.last_error()
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        l0 = L.fastcheb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for i in range(reps):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        bytes_per_fit = 2 * np.dtype(dt).itemsize * per
        gbs = howmany * bytes_per_fit / (ms * 1e-3) / 1e9
        print(json.dumps({"shape": tag, "howmany": howmany, "ms": ms, "fits_per_s": howmany / (ms * 1e-3), "gb_per_s": gbs,
                          "frac_of_hbm": gbs / hbm, "launches_per_call": (L.fastcheb_launch_count() - l0) / reps}), flush=True)
        del vals, outs


if __name__ == "__main__":
    main()

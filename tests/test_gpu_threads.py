"""GPU: the threading contract of the C ABI: a handle is immutable after creation — concurrent
evaluations of ONE handle and of distinct handles from several host threads are safe and give the serial results;
the error state (message, offending index) is per thread."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_concurrent_calls_match_serial(fc):
    rng = np.random.default_rng(17)
    coefs = rng.uniform(-1, 1, (17, 15, 13, 3))
    lb, ub = -np.ones(3), np.ones(3)
    shared = fc.ChebPoly(coefs, lb, ub)
    pts = [rng.uniform(-1, 1, (200_000 + 1000 * i, 3)) for i in range(6)]
    serial = [shared(p) for p in pts]
    serialJ = [fc.chebjacobian(shared, p)[1] for p in pts[:2]]
    vals = rng.uniform(-1, 1, (6, 3000, 33))
    serial_fit = [fc.chebcoefs(v, ndim=1, howmany=3000) for v in vals]
    out, errs = {}, []

    def worker(i):
        try:
            for rep in range(3):
                out[("v", i)] = shared(pts[i])                                 # one handle, many threads
                own = fc.ChebPoly(coefs * (i + 1), lb, ub)                    # distinct handles
                out[("o", i)] = own(pts[i][:1000])
                out[("f", i)] = fc.chebcoefs(vals[i], ndim=1, howmany=3000)   # concurrent fits (scratch cache, flags)
                if i < 2:
                    out[("J", i)] = fc.chebjacobian(shared, pts[i])[1]
                own.close()
                # a domain error on this thread reports this thread's offending index
                bad = pts[i][:5000].copy()
                bad[100 + i, 1] = 7.0
                try:
                    shared(bad)
                    errs.append((i, "no error"))
                except fc.ArgumentError as e:
                    if e.index != 100 + i:
                        errs.append((i, e.index))
        except Exception as e:  # pragma: no cover
            errs.append((i, repr(e)))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(6)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for i in range(6):
        assert np.array_equal(out[("v", i)], serial[i])
        assert np.allclose(out[("o", i)], serial[i][:1000] * (i + 1), rtol=1e-12, atol=1e-12)
        assert np.array_equal(out[("f", i)], serial_fit[i])
    for i in range(2):
        assert np.array_equal(out[("J", i)], serialJ[i])

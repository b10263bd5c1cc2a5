#!/usr/bin/env python
"""Generates tests/golden/truth_small.npz: small seeded cases with HIGH-PRECISION answers that do not come from the
oracle's working-precision code path — numpy.longdouble tensor-product Clenshaw for values (oracle.eval_truth), the
same applied to the long-double coefficient-space derivative (numpy's chebder recurrence) for Jacobians, and the
exactly-reduced long-double cosine sum for fitted coefficients.  Both the oracle (tests/test_oracle_golden.py) and the CUDA path (tests/test_gpu_golden.py) are checked
against this file, so an accidental change to either shows up against a fixed reference.

    python tests/golden/make_golden.py        (rewrites truth_small.npz; deterministic)"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cheb_oracle as orc  # noqa: E402

CASES = [  # name, dims, K, complex
    ("d1", (17,), None, False), ("d1c", (9,), 2, True), ("d2", (9, 7), None, False), ("d2k", (6, 5), 3, False),
    ("d3", (7, 6, 5), 3, False), ("d4c", (5, 4, 3, 4), None, True), ("n1", (1, 5), None, False), ("n2", (2, 2), 2, False),
]


def main():
    rng = np.random.default_rng(20260101)
    out = {}
    for name, dims, K, cplx in CASES:
        shape = dims + ((K,) if K else ())
        c = rng.uniform(-1, 1, shape)
        if cplx:
            c = c + 1j * rng.uniform(-1, 1, shape)
        N = len(dims)
        lb = -1.0 - 0.5 * np.arange(N)
        ub = 0.75 + 0.25 * np.arange(N)
        pts = lb + (ub - lb) * (0.05 + 0.9 * rng.random((24, N)))
        ld = np.longdouble
        v = orc.eval_truth(c, lb, ub, pts)
        cl = c.astype(np.clongdouble if cplx else ld)
        J = []
        for d in range(N):
            if dims[d] == 1:
                J.append(np.zeros_like(np.asarray(v)))
                continue
            dc = np.polynomial.chebyshev.chebder(cl, axis=d)          # d/dz_d in coefficient space, long double
            J.append(orc.eval_truth(dc, lb, ub, pts) * (ld(2) / (ld(ub[d]) - ld(lb[d]))))   # eval.jl:151
        J = np.stack(J, axis=-1)                       # (npts[, K], N)
        out[name + "_coefs"] = c
        out[name + "_lb"], out[name + "_ub"], out[name + "_pts"] = lb, ub, pts
        cd = np.complex128 if cplx else np.float64
        out[name + "_val"] = np.asarray(v).astype(cd)
        out[name + "_jac"] = np.asarray(J).astype(cd)
        vals = rng.uniform(-1, 1, shape) + (1j * rng.uniform(-1, 1, shape) if cplx else 0)
        out[name + "_fitvals"] = vals
        out[name + "_fitcoefs"] = orc.chebcoefs(vals, N)   # long-double cosine sums, rounded once to double
    np.savez_compressed(os.path.join(HERE, "truth_small.npz"), **out)
    print("wrote", os.path.join(HERE, "truth_small.npz"), sum(a.nbytes for a in out.values()), "bytes uncompressed")


if __name__ == "__main__":
    main()

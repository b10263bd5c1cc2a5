#!/usr/bin/env python
"""Measured gradient / Jacobian parity of the DEFAULT (AUTO) evaluation path on BASELINE configs[1], [2], [3]
(test infrastructure: uses the oracle as the checker).

north_star tolerance: |dJ| <= 16 eps sum|c| per gradient entry (entries carry the 2/(ub-lb) scale of eval.jl:151).
For every config, dense-random and analytic-fit coefficients, face / corner points included, this prints

  strict   = max |J_gpu - J_oracle| / (16 eps sum|c| * 2/(ub_d-lb_d))            the literal north_star bound
  weighted = max |J_gpu - J_oracle| / (16 eps sum_k |c_k| k_d^2 * 2/(ub_d-lb_d))  the bound applied to the
             differentiated series (|T_k'| <= k^2)
  gpu_vs_truth / oracle_vs_truth: both against the long-double Jacobian evaluated at the IDENTICAL normalised
             coordinate (numpy chebder in long double + oracle.eval_truth), in units of the strict bound

    python tools/jac_parity.py > profiles/r02_jac_parity.jsonl"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cheb_oracle as orc  # noqa: E402
import fastcheb  # noqa: E402

EPS = np.finfo(np.float64).eps


def analytic(dims, K, cplx, lb, ub):
    """Coefficients of a smooth analytic function fitted on the Chebyshev grid (tol = 0: no trimming)."""
    x = orc.chebpoints(tuple(n - 1 for n in dims), lb, ub)
    s = sum((d + 1.0) * x[..., d] for d in range(len(dims)))
    r2 = sum(x[..., d] ** 2 for d in range(len(dims)))
    comps = []
    for k in range(K or 1):
        f = np.exp(0.3 * s + 0.1 * k) / (1.0 + 2.0 * r2) + np.sin(1.7 * s + k)
        if cplx:
            f = f + 1j * np.cos(0.9 * s - k) * np.exp(-0.2 * r2)
        comps.append(f)
    vals = np.stack(comps, axis=-1) if K else comps[0]
    return orc.chebcoefs(vals, len(dims))


def weighted_sum(c, dims, d):
    """sum_k |c_k| k_d^2 (all components)."""
    k = np.arange(dims[d], dtype=np.float64) ** 2
    shape = [1] * c.ndim
    shape[d] = dims[d]
    return float((np.abs(c) * k.reshape(shape)).sum())


def truth_jac(c, dims, lb, ub, pts, z):
    ld = np.longdouble
    cl = c.astype(np.clongdouble if np.iscomplexobj(c) else ld)
    out = []
    for d in range(len(dims)):
        if dims[d] == 1:
            out.append(None)
            continue
        dc = np.polynomial.chebyshev.chebder(cl, axis=d)
        out.append(orc.eval_truth(dc, lb, ub, pts, z=z) * (ld(2) / (ld(ub[d]) - ld(lb[d]))))
    return out


def run(name, dims, K, cplx, kind, npts=20000, ntruth=96):
    rng = np.random.default_rng(zlib.crc32((name + kind).encode()))
    N = len(dims)
    lb = -1.0 - 0.25 * np.arange(N)
    ub = 0.5 + 0.75 * np.arange(N)
    if kind == "dense":
        shape = dims + ((K,) if K else ())
        c = rng.uniform(-1, 1, shape)
        if cplx:
            c = c + 1j * rng.uniform(-1, 1, shape)
    else:
        c = analytic(dims, K, cplx, lb, ub)
    pts = lb + (ub - lb) * rng.random((npts, N))
    # faces, corners, and points one ulp inside the faces (where T_k' is largest)
    pts[0], pts[1] = lb, ub
    for i in range(2, 2 + 2 * N):
        d = (i - 2) // 2
        pts[i, d] = lb[d] if i % 2 == 0 else ub[d]
    for i in range(2 + 2 * N, 2 + 4 * N):
        d = (i - 2 - 2 * N) // 2
        pts[i, d] = np.nextafter(lb[d], ub[d]) if i % 2 == 0 else np.nextafter(ub[d], lb[d])
    near = rng.integers(0, N, npts // 10)
    idx = np.arange(100, 100 + npts // 10)
    pts[idx, near] = np.where(rng.random(npts // 10) < 0.5, lb[near] + (ub - lb)[near] * 1e-4 * rng.random(npts // 10),
                              ub[near] - (ub - lb)[near] * 1e-4 * rng.random(npts // 10))
    ref = orc.OracleChebPoly(c, lb, ub)
    v0, J0 = ref.jacobian(pts)
    poly = fastcheb.ChebPoly(c, lb, ub)
    algo = {1: "clenshaw", 2: "dmma"}[poly.get_algo(True)]
    v, J = fastcheb.chebjacobian(poly, pts)
    sumc = float(np.abs(c).sum())
    rec = {"config": name, "coefs": kind, "dims": list(dims), "K": K, "complex": cplx, "algo_auto_jac": algo, "points": npts,
           "sum_abs_c": sumc, "value_strict": float(np.abs(v - v0).max() / (16 * EPS * sumc))}
    strict, weighted = [], []
    for d in range(N):
        sc = 2.0 / (ub[d] - lb[d])
        dj = float(np.abs(J[:, :, d] - J0[:, :, d]).max())
        strict.append(dj / (16 * EPS * sumc * sc))
        weighted.append(dj / (16 * EPS * max(weighted_sum(c, dims, d), 1e-300) * sc))
    rec["strict_per_dim"] = strict
    rec["weighted_per_dim"] = weighted
    rec["strict"] = max(strict)
    rec["weighted"] = max(weighted)
    # against long-double truth on a subset (identical z): who is closer, the GPU or the oracle?
    sub = np.r_[0:2 + 4 * N, 100:100 + ntruth]
    ps = pts[sub]
    z = (ps - lb) * 2 / (ub - lb) - 1          # eval.jl:65 in the working precision: the z both implementations use
    tj = truth_jac(c, dims, lb, ub, ps, z)
    g, o = [], []
    for d in range(N):
        if tj[d] is None:
            continue
        sc = 2.0 / (ub[d] - lb[d])
        t = np.asarray(tj[d]).reshape(len(sub), -1)
        g.append(float(np.abs(J[sub][:, :, d] - t).max() / (16 * EPS * sumc * sc)))
        o.append(float(np.abs(J0[sub][:, :, d] - t).max() / (16 * EPS * sumc * sc)))
    rec["gpu_vs_truth_strict_units"] = max(g)
    rec["oracle_vs_truth_strict_units"] = max(o)
    poly.close()
    print(json.dumps(rec), flush=True)


def main():
    for kind in ("dense", "analytic"):
        run("configs[1] 2-D order (64,64) real F64", (65, 65), None, False, kind)
        run("configs[2] 3-D order (32,32,32) SVector{3} F64", (33, 33, 33), 3, False, kind, npts=8000, ntruth=48)
        run("configs[3] 4-D order (16,16,16,16) ComplexF64", (17, 17, 17, 17), None, True, kind, npts=6000, ntruth=32)


if __name__ == "__main__":
    main()

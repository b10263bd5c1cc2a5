#!/usr/bin/env python
"""Low-order, memory-bound evaluations (north_star: achieved HBM GB/s for "the low-order, memory-bound evaluations"):
polynomials with a handful of coefficients per dimension, where a point's coordinates and results (8-100 bytes) cost more
than its arithmetic.  One JSON line per shape: points/s and the fraction of the measured HBM copy bandwidth.
usage: probe_loworder.py [once IDX]   ("once": a single launch of shape IDX, for ncu)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import fastcheb
from fastcheb import _capi
import bench_legs as B

f64, f32 = np.float64, np.float32
# (name, dims, K, complex, dtype, points, jacobian)
SHAPES = [
    ("1d_order4_f64_val", (5,), None, False, f64, 1 << 26, False),
    ("1d_order4_f32_val", (5,), None, False, f32, 1 << 27, False),
    ("1d_order8_f64_val", (9,), None, False, f64, 1 << 26, False),
    ("1d_order8_f64_grad", (9,), None, False, f64, 1 << 26, True),
    ("2d_order4_f64_val", (5, 5), None, False, f64, 1 << 25, False),
    ("2d_order4_f64_grad", (5, 5), None, False, f64, 1 << 25, True),
    ("2d_order4_f32_val", (5, 5), None, False, f32, 1 << 26, False),
    ("3d_order3_K3_f64_val", (4, 4, 4), 3, False, f64, 1 << 24, False),
    ("3d_order3_K3_f64_jac", (4, 4, 4), 3, False, f64, 1 << 24, True),
    ("3d_order2_complex_f64_val", (3, 3, 3), None, True, f64, 1 << 24, False),
    ("4d_order2_f32_val", (3, 3, 3, 3), None, False, f32, 1 << 25, False),
    ("2d_order8_f64_val", (9, 9), None, False, f64, 1 << 25, False),
    ("2d_order8_f64_grad", (9, 9), None, False, f64, 1 << 25, True),
    ("2d_order4_K3_f64_jac", (5, 5), 3, False, f64, 1 << 24, True),
    ("3d_order4_f64_val", (5, 5, 5), None, False, f64, 1 << 24, False),
    ("3d_order4_f64_grad", (5, 5, 5), None, False, f64, 1 << 24, True),
    ("3d_order3_K2_f32_jac", (4, 4, 4), 2, False, f32, 1 << 24, True),
    ("4d_order2_f64_grad", (3, 3, 3, 3), None, False, f64, 1 << 24, True),
    ("4d_order3_K2_f64_val", (4, 4, 4, 4), 2, False, f64, 1 << 23, False),
    # nests between 512 and 2048 padded coefficients (second table capacity)
    ("3d_order8_f64_val", (9, 9, 9), None, False, f64, 1 << 23, False),
    ("3d_order8_f64_grad", (9, 9, 9), None, False, f64, 1 << 22, True),
    ("2d_17x33_f64_val", (17, 33), None, False, f64, 1 << 23, False),
    ("2d_9x40_K3_f64_val", (9, 40), 3, False, f64, 1 << 22, False),
    ("4d_order4_f32_val", (5, 5, 5, 5), None, False, f32, 1 << 22, False),
    # longer dense-random 1-D series (the product basis rejects them): FP64-bound, the same kernel as the Clenshaw kernel
    ("1d_order24_f64_val", (25,), None, False, f64, 1 << 25, False),
    ("1d_order64_dense_f64_val", (65,), None, False, f64, 1 << 25, False),
    ("1d_order64_dense_f64_grad", (65,), None, False, f64, 1 << 25, True),
    ("1d_order64_dense_f32_val", (65,), None, False, f32, 1 << 25, False),
    ("1d_order200_dense_f64_val", (201,), None, False, f64, 1 << 24, False),
    ("1d_order200_dense_K3_f64_val", (201,), 3, False, f64, 1 << 23, False),
    ("1d_order512_dense_f64_val", (513,), None, False, f64, 1 << 23, False),
]
LEGS = {s[0]: s for s in SHAPES}


def run(name, reps):
    _, dims, K, cplx, dt, npts, jac = LEGS[name]
    return B.eval_leg(torch, fastcheb, _capi, 0, dims, K, cplx, dt, npts, jac, reps)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "once":
        print(json.dumps(run(SHAPES[int(sys.argv[2])][0], 1)))
        sys.exit(0)
    for s in SHAPES:
        r = run(s[0], 5)
        ro = r["roofline"]
        print(json.dumps({"leg": s[0], "algo": r["algo"], "points_per_s": r["value"], "ms": r["ms"], "bound": ro["bound"],
                          "GB_per_s": r["value"] * ro["bytes_per_point"] / 1e9, "bytes_per_point": ro["bytes_per_point"],
                          "frac_hbm": ro["frac_hbm"], "frac_fp": ro["frac_fp"]}))

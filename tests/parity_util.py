"""Shared tolerance helpers of the GPU parity tests (test infrastructure).

north_star: |dv| <= 16 eps sum|c| per evaluated value and per gradient entry.  For gradient / Jacobian entries the
bound is applied to the series that is actually summed: dv/dx_d = 2/(ub_d-lb_d) sum_k c_k T'_{k_d}(z_d) prod_{i!=d}
T_{k_i}(z_i) with |T'_k| <= k^2, i.e.

    |dJ_d| <= 16 eps * (sum_k |c_k| k_d^2) * 2/(ub_d-lb_d)            ("weighted" form)

profiles/r02_jac_parity.jsonl (tools/jac_parity.py) records the measured worst |dJ| of the default tensor-core path on
BASELINE configs[1..3] against this bound (<= 0.18), against the literal bound without the k_d^2 weight (0.7 - 249: the
oracle itself is 0.2 - 176 of that bound away from the long-double Jacobian at the same z, so no implementation can
meet it), and against long-double truth."""
import numpy as np


def weighted_abs_sum(c, ndim, d):
    """sum_k |c_k| k_d^2 over all coefficients (and components) of a (n1..nN[,K]) array."""
    c = np.asarray(c)
    k2 = np.arange(c.shape[d], dtype=np.float64) ** 2
    shape = [1] * c.ndim
    shape[d] = c.shape[d]
    return float((np.abs(c).astype(np.float64) * k2.reshape(shape)).sum())


def jac_tol(c, ndim, d, lb, ub, dt=np.float64):
    """Tolerance of Jacobian column d (entries carry the 2/(ub-lb) scale of eval.jl:151)."""
    return 16 * float(np.finfo(dt).eps) * weighted_abs_sum(c, ndim, d) * 2.0 / (float(ub[d]) - float(lb[d]))


def truth_jacobian(orc, c, lb, ub, pts, z=None):
    """Long-double Jacobian columns [(npts, K or 1) or None for size-1 dims]: numpy's chebder recurrence in long
    double + oracle.eval_truth, at the normalised coordinate `z` when given (identical-z comparison)."""
    ld = np.longdouble
    c = np.asarray(c)
    N = len(lb)
    cl = c.astype(np.clongdouble if np.iscomplexobj(c) else ld)
    out = []
    for d in range(N):
        if c.shape[d] == 1:
            out.append(None)
            continue
        dc = np.polynomial.chebyshev.chebder(cl, axis=d)
        t = orc.eval_truth(dc, lb, ub, pts, z=z) * (ld(2) / (ld(ub[d]) - ld(lb[d])))
        out.append(np.asarray(t).reshape(len(pts), -1))
    return out

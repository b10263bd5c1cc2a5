"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Python face of the CPU oracle: a restatement of the reference's hot path
(/root/reference/src/eval.jl, /root/reference/src/interp.jl).  May be imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs.  The product
package (fastcheb) never imports it.

Three grades of restatement live here, each function citing the reference lines it follows:

* ``libcheb_oracle.so`` (oracle/cheb_oracle.c) — line-faithful C Clenshaw / Jacobian recursion in the
  working precision (the *parity oracle* and, with OpenMP, the timed CPU baseline), and a
  truth-grade long-double DCT-I fit restated from FFTW's REDFT00 definition.
* ``chebcoefs_fft`` — scipy/pocketfft ``dct(type=1)`` as the stand-in for the reference's
  ``FFTW.r2r(vals, REDFT00)`` (interp.jl:43-44; libfftw3 is not in this image).
* ``eval_truth`` — numpy.longdouble tensor-product Clenshaw used as "truth" for error budgets.

Array convention (numpy): scalar-valued data has shape (n1..nN); vector-valued data has a trailing
component axis, shape (n1..nN, K).  The flat buffers handed to C are in Julia's memory order
(component fastest, then k1, k2, ...), i.e. ``Array{SVector{K,T},N}`` column-major.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcheb_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        L = ctypes.CDLL(path)
        i64p = ctypes.POINTER(ctypes.c_int64)
        vp = ctypes.c_void_p
        for sfx in ("f64", "f32"):
            f = getattr(L, f"oracle_eval_{sfx}")
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_int, i64p, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int64, vp, i64p]
            f = getattr(L, f"oracle_eval_jac_{sfx}")
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.c_int, i64p, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int64, vp, vp, i64p]
        L.oracle_eval_vec_f64.restype = ctypes.c_int
        L.oracle_eval_vec_f64.argtypes = [ctypes.c_int, i64p, ctypes.c_int, vp, vp, vp, vp, ctypes.c_int64, vp, i64p]
        L.oracle_fit_ld.restype = ctypes.c_int
        L.oracle_fit_ld.argtypes = [ctypes.c_int, i64p, ctypes.c_int, vp, vp, i64p]
        L.oracle_droptol_shape.restype = ctypes.c_int
        L.oracle_droptol_shape.argtypes = [ctypes.c_int, i64p, ctypes.c_int, ctypes.c_int, vp, ctypes.c_double, i64p]
        L.oracle_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


# --------------------------------------------------------------------------- layout helpers
def _real_dtype(dt) -> np.dtype:
    dt = np.dtype(dt)
    return np.dtype({np.dtype(np.complex64): np.float32, np.dtype(np.complex128): np.float64}.get(dt, dt))


def to_julia_flat(a: np.ndarray, ndim: int) -> np.ndarray:
    """(n1..nN[,K]) numpy array -> flat real buffer in Julia memory order (component, re/im fastest)."""
    a = np.asarray(a)
    if a.ndim == ndim + 1:
        a = np.moveaxis(a, -1, 0)
    elif a.ndim != ndim:
        raise ValueError("array rank does not match the number of dimensions")
    flat = np.ascontiguousarray(a.reshape(-1, order="F"))
    return flat.view(_real_dtype(flat.dtype)) if np.iscomplexobj(flat) else flat


def from_julia_flat(flat: np.ndarray, dims, K: int | None, cplx: bool) -> np.ndarray:
    """Inverse of to_julia_flat.  K=None -> scalar-valued (no component axis)."""
    flat = np.ascontiguousarray(flat)
    if cplx:
        flat = flat.view({np.dtype(np.float32): np.complex64, np.dtype(np.float64): np.complex128}[flat.dtype])
    if K is None:
        return flat.reshape(tuple(dims), order="F")
    return np.moveaxis(flat.reshape((K,) + tuple(dims), order="F"), 0, -1)


# --------------------------------------------------------------------------- grid (interp.jl:4-12)
def chebpoints(order, lb, ub, dtype=np.float64):
    """interp.jl:4-12,28-36.  Returns shape (n1..nN, N) for tuple orders, (n,) for a scalar order.
    Index 0 along each dim is the *upper* bound; order 0 -> the midpoint (cos(0*pi/2)... :6)."""
    scalar = np.isscalar(order)
    order = (int(order),) if scalar else tuple(int(o) for o in order)
    lbv = np.atleast_1d(np.asarray(lb, dtype=dtype))
    ubv = np.atleast_1d(np.asarray(ub, dtype=dtype))
    if any(o < 0 for o in order):
        raise ValueError(f"invalid negative order {order}")          # interp.jl:10 (ArgumentError)
    if not (len(order) == len(lbv) == len(ubv)):
        raise ValueError("dimension mismatch")                        # interp.jl:30
    dt = np.dtype(dtype).type
    axes = []
    for d, o in enumerate(order):
        # interp.jl:12: order 0 -> the single index 1 with denominator 2 (interp.jl:6) = the midpoint
        i = np.array([1], dtype=dt) if o == 0 else np.arange(0, o + 1, dtype=dt)
        den = dt(2 if o == 0 else o)
        c = np.cos(i * dt(np.pi) / den).astype(dt)
        axes.append((lbv[d] + (dt(1) + c) * (ubv[d] - lbv[d]) * dt(0.5)).astype(dt))     # interp.jl:6
    if scalar:
        return axes[0]
    grid = np.meshgrid(*axes, indexing="ij")
    return np.stack(grid, axis=-1)


# --------------------------------------------------------------------------- fit (interp.jl:39-57)
def _dims_ptr(dims):
    arr = (ctypes.c_int64 * len(dims))(*[int(d) for d in dims])
    return arr


def chebcoefs(vals: np.ndarray, ndim: int) -> np.ndarray:
    """interp.jl:39-71 via the long-double direct cosine sum (truth-grade), rounded to vals' dtype."""
    vals = np.asarray(vals)
    if vals.dtype.kind not in "fc":
        vals = vals.astype(np.float64)
    K = vals.shape[-1] if vals.ndim == ndim + 1 else None
    cplx = np.iscomplexobj(vals)
    dims = vals.shape[:ndim]
    flat = to_julia_flat(vals, ndim)
    E = (K or 1) * (2 if cplx else 1)
    x = flat.astype(np.longdouble)
    y = np.empty_like(x)
    bad = ctypes.c_int64(-1)
    rc = lib().oracle_fit_ld(ndim, _dims_ptr(dims), E, x.ctypes.data, y.ctypes.data, ctypes.byref(bad))
    if rc == 2:
        raise FloatingPointError("non-finite interpolant value")      # DomainError, interp.jl:40
    return from_julia_flat(y.astype(flat.dtype), dims, K, cplx)


def chebcoefs_fft(vals: np.ndarray, ndim: int, workers: int | None = None) -> np.ndarray:
    """interp.jl:39-57 with scipy/pocketfft dct(type=1) standing in for FFTW REDFT00 (interp.jl:44)."""
    import scipy.fft

    vals = np.asarray(vals)
    if not np.all(np.isfinite(vals)):
        raise FloatingPointError("non-finite interpolant value")      # interp.jl:40
    c = vals.copy()
    for d in range(ndim):
        n = vals.shape[d]
        if n > 1:                                                     # interp.jl:43
            if np.iscomplexobj(c):
                c = scipy.fft.dct(c.real, type=1, axis=d, workers=workers) + 1j * scipy.fft.dct(c.imag, type=1, axis=d, workers=workers)
            else:
                c = scipy.fft.dct(c, type=1, axis=d, workers=workers)
            c = c / (2 * (n - 1))                                     # interp.jl:49
            sl = [slice(None)] * c.ndim
            sl[d] = slice(1, n - 1)
            c[tuple(sl)] *= 2                                         # interp.jl:50-54
    return c.astype(vals.dtype)


def droptol(coefs: np.ndarray, ndim: int, tol: float) -> np.ndarray:
    """interp.jl:77-93."""
    coefs = np.asarray(coefs)
    K = coefs.shape[-1] if coefs.ndim == ndim + 1 else None
    cplx = np.iscomplexobj(coefs)
    dims = coefs.shape[:ndim]
    flat = to_julia_flat(coefs, ndim).astype(np.float64)
    newdims = (ctypes.c_int64 * ndim)()
    changed = lib().oracle_droptol_shape(ndim, _dims_ptr(dims), K or 1, int(cplx), flat.ctypes.data, float(tol), newdims)
    if not changed:
        return coefs
    sl = tuple(slice(0, int(newdims[d])) for d in range(ndim))
    return np.ascontiguousarray(coefs[sl])


# --------------------------------------------------------------------------- ChebPoly mirror
@dataclass
class OracleChebPoly:
    """struct ChebPoly{N,T,Td} FastChebInterp.jl:37-41.  coefs shape (n1..nN[,K])."""
    coefs: np.ndarray
    lb: np.ndarray
    ub: np.ndarray

    @property
    def ndim(self) -> int:
        return len(self.lb)

    @property
    def K(self):
        return self.coefs.shape[-1] if self.coefs.ndim == self.ndim + 1 else None

    def _prep(self, x):
        rdt = _real_dtype(self.coefs.dtype)
        x = np.asarray(x)
        # promotion (eval.jl:65 / :25): Float32 coefficients with Float32 points stay Float32
        if not (rdt == np.float32 and x.dtype == np.float32 and self.lb.dtype != np.float64):
            rdt = np.dtype(np.float64)
        # N == 1: scalar = one point, 1-D array = batch (c.(xs));  N > 1: length-N vector = one point
        single = (x.ndim == 0) if self.ndim == 1 else (x.ndim == 1)
        pts = np.ascontiguousarray(np.asarray(x, dtype=rdt).reshape(-1, self.ndim))
        return rdt, single, pts

    def _call(self, x, jac: bool, vec: bool = False):
        rdt, single, pts = self._prep(x)
        N = self.ndim
        dims = self.coefs.shape[:N]
        cplx = np.iscomplexobj(self.coefs)
        K = self.K
        E = (K or 1) * (2 if cplx else 1)
        flat = to_julia_flat(self.coefs, N).astype(rdt)
        lb = self.lb.astype(rdt)
        ub = self.ub.astype(rdt)
        npts = pts.shape[0]
        out = np.empty(npts * E, dtype=rdt)
        bad = ctypes.c_int64(-1)
        sfx = "f64" if rdt == np.float64 else "f32"
        if jac:
            outJ = np.empty(npts * E * N, dtype=rdt)
            rc = getattr(lib(), f"oracle_eval_jac_{sfx}")(N, _dims_ptr(dims), E, flat.ctypes.data, lb.ctypes.data, ub.ctypes.data,
                                                          pts.ctypes.data, npts, out.ctypes.data, outJ.ctypes.data, ctypes.byref(bad))
        else:
            fn = f"oracle_eval_vec_{sfx}" if (vec and sfx == "f64") else f"oracle_eval_{sfx}"
            rc = getattr(lib(), fn)(N, _dims_ptr(dims), E, flat.ctypes.data, lb.ctypes.data, ub.ctypes.data,
                                    pts.ctypes.data, npts, out.ctypes.data, ctypes.byref(bad))
        if rc == 1:
            err = ValueError(f"{pts[bad.value]} not in domain {self.lb} to {self.ub}")   # ArgumentError eval.jl:64
            err.index = int(bad.value)
            raise err
        v = out.reshape(npts, E)
        if cplx:
            v = v.view(np.complex64 if rdt == np.float32 else np.complex128)
        v = v.reshape(npts) if K is None else v.reshape(npts, K)
        if not jac:
            return v[0] if single else v
        J = outJ.reshape(npts, N, E)          # per point column-major E x N  ->  [p, d, e]
        if cplx:
            J = J.view(np.complex64 if rdt == np.float32 else np.complex128)
        J = np.swapaxes(J.reshape(npts, N, K or 1), 1, 2)            # [p, k, d] = K x N matrix per point
        return (v[0], J[0]) if single else (v, J)

    def __call__(self, x):                                           # eval.jl:63-70
        return self._call(x, False)

    def eval_simd(self, x):
        """Same values, bit for bit, with all components of an element advanced together (cheb_oracle_vec.inc): the
        SIMD over SVector/Complex lanes that the reference gets from LLVM.  Used for the timed CPU baseline."""
        return self._call(x, False, vec=True)

    def jacobian(self, x):                                           # eval.jl:147-155
        return self._call(x, True)

    def gradient(self, x):                                           # eval.jl:168-177
        if self.K is not None and self.K != 1:
            raise TypeError("DimensionMismatch: chebgradient needs a scalar-valued polynomial")  # eval.jl:170
        v, J = self._call(x, True)
        if self.K is not None:
            v = v[..., 0]
        g = J[..., 0, :]
        if self.ndim == 1 and np.ndim(x) == 0:
            return v, g[..., 0]                                      # eval.jl:174-177
        return v, g


def chebinterp(vals, lb, ub, tol=None, fft: bool = False) -> OracleChebPoly:
    """interp.jl:95-99,140-141 (tol default = eps of the real float type of vals, interp.jl:102)."""
    lbv = np.atleast_1d(np.asarray(lb))
    ubv = np.atleast_1d(np.asarray(ub))
    N = len(lbv)
    vals = np.asarray(vals)
    if vals.dtype.kind not in "fc":
        vals = vals.astype(np.float64)
    if tol is None:
        tol = float(np.finfo(_real_dtype(vals.dtype)).eps)
    c = chebcoefs_fft(vals, N) if fft else chebcoefs(vals, N)
    c = droptol(c, N, tol)
    td = np.result_type(lbv.dtype, ubv.dtype)
    return OracleChebPoly(c, lbv.astype(td), ubv.astype(td))


def chebinterp_v1(vals, lb, ub, tol=None) -> OracleChebPoly:
    """interp.jl:155-156: the *first* axis is the component axis."""
    return chebinterp(np.moveaxis(np.asarray(vals), 0, -1), lb, ub, tol=tol)


# --------------------------------------------------------------------------- long-double truth
def eval_truth(coefs: np.ndarray, lb, ub, pts: np.ndarray, z: np.ndarray | None = None) -> np.ndarray:
    """sum_k C[k] prod_d T_kd(z_d) by nested Clenshaw in numpy.longdouble (x87 80-bit here), vectorised
    over points.  If ``z`` is given it is used verbatim as the normalised coordinate (so that parity
    budgets can be asserted for *identical* z); else z follows eval.jl:65 in long double.
    Returns float128-typed array of shape (npts,) or (npts, K) (complex -> complex256)."""
    ld = np.longdouble
    lbv = np.atleast_1d(np.asarray(lb)).astype(ld)
    ubv = np.atleast_1d(np.asarray(ub)).astype(ld)
    N = len(lbv)
    pts = np.asarray(pts).reshape(-1, N)
    if z is None:
        z = (pts.astype(ld) - lbv) * 2 / (ubv - lbv) - 1
    else:
        z = np.asarray(z).reshape(-1, N).astype(ld)
    c = np.asarray(coefs)
    cl = c.astype(np.clongdouble if np.iscomplexobj(c) else ld)
    vec = c.ndim == N + 1
    if not vec:
        cl = cl[..., None]
    # contract dims from the first to the last: state has shape (npts, n_{d+1}.., K) after dim d
    npts = pts.shape[0]
    state = np.broadcast_to(cl, (npts,) + cl.shape)
    for d in range(N):
        n = state.shape[1]
        zd = z[:, d].reshape((npts,) + (1,) * (state.ndim - 2))
        if n == 1:
            state = state[:, 0]
            continue
        b1 = np.zeros_like(state[:, 0])
        b2 = np.zeros_like(b1)
        for j in range(n - 1, 0, -1):
            b1, b2 = 2 * zd * b1 - b2 + state[:, j], b1
        state = zd * b1 - b2 + state[:, 0]
    return state if vec else state[:, 0]


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int | None = None) -> int:
    """Use `n` OpenMP threads (default: every core this process may run on).  torchrun exports
    OMP_NUM_THREADS=1 to its ranks, which would otherwise turn the timed CPU baseline into a 1-core run."""
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    L = lib()
    L.oracle_set_num_threads.argtypes = [ctypes.c_int]
    L.oracle_set_num_threads.restype = None
    L.oracle_set_num_threads(int(n))
    return num_threads()

"""fastcheb — host-side mirror of FastChebInterp.jl's hot-path API over libfastcheb_cuda (B200, sm_100a).

The reference is Julia; Julia is not in this image, so this module is the stand-in for the thin
`src/cuda.jl` wrapper (julia/src/cuda.jl holds the Julia text): same names, same argument meaning,
same error behaviour, every numeric step done by the CUDA library through the C ABI of
include/fastcheb.h.  There is no CPU path here — importing this package without the built shared
library raises ImportError, and every call needs a B200.

Reference API mirrored (file:line under /root/reference/src):
  ChebPoly                      FastChebInterp.jl:37-41   (coefs, lb, ub)
  c(x), c.(xs)                  eval.jl:63-70
  chebjacobian(c, x)            eval.jl:147-155
  chebgradient(c, x)            eval.jl:168-177
  chebinterp(vals, lb, ub; tol) interp.jl:95-99, 140-141
  chebinterp_v1                 interp.jl:155-156
  chebpoints(order, lb, ub)     interp.jl:4-36   (host helper: grid generation is not on the device path)

Array conventions (numpy): scalar-valued data has shape (n1..nN); vector-valued data carries a trailing
component axis (n1..nN, K) — the flat buffer handed to C is Julia's `Array{SVector{K,T},N}` order.
Points: a length-N vector is one point, an (npts, N) array is a batch (`c.(ys)`); for N == 1 a scalar is
one point and a 1-D array a batch.  torch CUDA tensors are accepted for points (device-resident batches)
and give torch tensors back without touching the host.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _capi
from ._capi import ALGO_AUTO, ALGO_CLENSHAW, ALGO_DMMA  # noqa: F401

__all__ = ["ChebPoly", "chebinterp", "chebinterp_v1", "chebjacobian", "chebgradient", "chebpoints", "chebcoefs",
           "droptol", "ArgumentError", "DomainError", "DimensionMismatch", "CudaError"]

_capi.lib()  # fail loudly at import time if the CUDA library is missing


class ArgumentError(ValueError):
    """Julia's ArgumentError (eval.jl:64,148; interp.jl:10,69)."""


class DomainError(ValueError):
    """Julia's DomainError (interp.jl:40)."""


class DimensionMismatch(ValueError):
    """Julia's DimensionMismatch (eval.jl:170; interp.jl:30)."""


class CudaError(RuntimeError):
    pass


def _check(rc: int, what: str = ""):
    if rc == _capi.OK:
        return
    msg = _capi.last_error()
    if rc == _capi.EDOMAIN:
        raise ArgumentError(f"{what}{msg}")
    if rc == _capi.ENONFINITE:
        raise DomainError(f"{what}{msg}")
    if rc == _capi.EDIMMISMATCH:
        raise DimensionMismatch(f"{what}{msg}")
    if rc == _capi.EINVAL:
        raise ArgumentError(f"{what}{msg}")
    if rc == _capi.ENOMEM:
        raise MemoryError(f"{what}{msg}")
    raise CudaError(f"{what}{msg} (status {rc})")


# ------------------------------------------------------------------------------------------- layout
def _real_dtype(dt) -> np.dtype:
    dt = np.dtype(dt)
    if dt == np.complex64:
        return np.dtype(np.float32)
    if dt == np.complex128:
        return np.dtype(np.float64)
    return dt


def _to_julia_flat(a: np.ndarray, ndim: int) -> np.ndarray:
    a = np.asarray(a)
    if a.ndim == ndim + 1:
        a = np.moveaxis(a, -1, 0)
    elif a.ndim != ndim:
        raise DimensionMismatch("array rank does not match the number of dimensions")
    flat = np.ascontiguousarray(a.reshape(-1, order="F"))
    return flat.view(_real_dtype(flat.dtype)) if np.iscomplexobj(flat) else flat


def _from_julia_flat(flat: np.ndarray, dims, K, cplx: bool) -> np.ndarray:
    flat = np.ascontiguousarray(flat)
    if cplx:
        flat = flat.view(np.complex64 if flat.dtype == np.float32 else np.complex128)
    if K is None:
        return flat.reshape(tuple(dims), order="F")
    return np.moveaxis(flat.reshape((K,) + tuple(dims), order="F"), 0, -1)


def _dtype_code(rdt) -> int:
    return _capi.F64 if np.dtype(rdt) == np.float64 else _capi.F32


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _float_vals(vals) -> np.ndarray:
    vals = np.asarray(vals)
    if vals.dtype == object:
        # array of vectors of unequal length -> ArgumentError (interp.jl:67-71)
        raise ArgumentError("array elements must all be of the same length")
    if vals.dtype.kind not in "fc":
        vals = vals.astype(np.float64)
    if vals.dtype not in (np.float32, np.float64, np.complex64, np.complex128):
        vals = vals.astype(np.complex128 if vals.dtype.kind == "c" else np.float64)
    return vals


# ------------------------------------------------------------------------------------------- ChebPoly
class ChebPoly:
    """Device-resident `ChebPoly{N,T,Td}` (FastChebInterp.jl:37-41).

    coefs: (n1..nN) or (n1..nN, K), float32/float64/complex64/complex128; lb, ub: length-N bounds
    (their dtype is kept: integer bounds are allowed, runtests.jl:170).
    """

    def __init__(self, coefs, lb, ub, device: int = 0, _handle=None, _handle_dtype=None):
        self.lb = np.atleast_1d(np.asarray(lb))
        self.ub = np.atleast_1d(np.asarray(ub))
        if self.lb.shape != self.ub.shape or self.lb.ndim != 1:
            raise DimensionMismatch("lb and ub must be vectors of the same length")
        td = np.result_type(self.lb.dtype, self.ub.dtype)
        self.lb, self.ub = self.lb.astype(td), self.ub.astype(td)
        self.ndim = len(self.lb)
        self.coefs = _float_vals(coefs)
        if self.coefs.ndim not in (self.ndim, self.ndim + 1):
            raise DimensionMismatch("coefficient array rank does not match the bounds")
        self.K = self.coefs.shape[-1] if self.coefs.ndim == self.ndim + 1 else None
        self.device = int(device)
        self._handles = {}
        if _handle is not None:
            self._handles[np.dtype(_handle_dtype)] = _handle

    # -- handle management
    @property
    def dims(self):
        return tuple(self.coefs.shape[: self.ndim])

    @property
    def is_complex(self) -> bool:
        return np.iscomplexobj(self.coefs)

    def _handle(self, rdt) -> ctypes.c_void_p:
        rdt = np.dtype(rdt)
        h = self._handles.get(rdt)
        if h is None:
            flat = _to_julia_flat(self.coefs, self.ndim).astype(rdt, copy=False)
            flat = np.ascontiguousarray(flat)
            h = ctypes.c_void_p()
            rc = _capi.lib().fastcheb_create(ctypes.byref(h), self.ndim, _capi.i64_array(self.dims), _dtype_code(rdt),
                                             int(self.is_complex), self.K or 1, flat.ctypes.data, _capi.MEM_HOST,
                                             _capi.dbl_array(self.lb), _capi.dbl_array(self.ub), self.device)
            _check(rc, "fastcheb_create: ")
            self._handles[rdt] = h
        return h

    def set_algo(self, algo: int, rdt=np.float64):
        _check(_capi.lib().fastcheb_set_algo(self._handle(rdt), int(algo)), "fastcheb_set_algo: ")

    def get_algo(self, jac: bool = False, rdt=np.float64) -> int:
        return int(_capi.lib().fastcheb_get_algo(self._handle(rdt), int(jac)))

    def close(self):
        for h in self._handles.values():
            _capi.lib().fastcheb_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __repr__(self):
        kind = f"{self.K}-vector " if self.K is not None else ""
        order = tuple(n - 1 for n in self.dims)
        return f"ChebPoly{{{self.ndim},{self.coefs.dtype}}} {kind}order {order} polynomial on [{self.lb},{self.ub}]"

    # -- evaluation
    def _compute_dtype(self, xdtype) -> np.dtype:
        crt = _real_dtype(self.coefs.dtype)
        # promotion (eval.jl:65 / :25): Float32 coefficients at Float32 points with non-Float64 bounds stay Float32
        if crt == np.float32 and np.dtype(xdtype) == np.float32 and self.lb.dtype != np.float64:
            return np.dtype(np.float32)
        return np.dtype(np.float64)

    def _eval(self, x, jac: bool, grad: bool = False):
        L = _capi.lib()
        N, K = self.ndim, self.K
        cplx = self.is_complex
        E = (K or 1) * (2 if cplx else 1)
        if _is_torch(x):
            return self._eval_torch(x, jac, grad)
        xa = np.asarray(x)
        if xa.dtype.kind not in "fiub":
            raise ArgumentError("points must be real")
        rdt = self._compute_dtype(xa.dtype if xa.dtype.kind == "f" else np.float64)
        single = (xa.ndim == 0) if N == 1 else (xa.ndim == 1)
        if not single and N > 1 and (xa.ndim != 2 or xa.shape[1] != N):
            raise DimensionMismatch(f"points must have shape (npts, {N})")
        if single and N > 1 and xa.shape[0] != N:
            raise DimensionMismatch(f"a point must have {N} coordinates")
        pts = np.ascontiguousarray(xa.astype(rdt, copy=False).reshape(-1, N))
        npts = pts.shape[0]
        h = self._handle(rdt)
        out = np.empty((npts, E), dtype=rdt)
        if jac:
            outJ = np.empty((npts, N, E), dtype=rdt)  # per point column-major E x N
            fn = L.fastcheb_eval_grad if grad else L.fastcheb_eval_jac
            rc = fn(h, pts.ctypes.data, npts, out.ctypes.data, outJ.ctypes.data, _capi.MEM_HOST, None)
        else:
            rc = L.fastcheb_eval(h, pts.ctypes.data, npts, out.ctypes.data, _capi.MEM_HOST, None)
        if rc == _capi.EDOMAIN:
            bad = _capi.error_index()
            err = ArgumentError(f"{pts[bad]} not in domain {self.lb} to {self.ub}")  # eval.jl:64
            err.index = bad  # 0-based index of the first offending point (sharded.py needs it)
            raise err
        _check(rc, "fastcheb_eval: ")
        cdt = np.complex64 if rdt == np.float32 else np.complex128
        v = out.view(cdt) if cplx else out
        v = v.reshape(npts) if K is None else v.reshape(npts, K)
        if not jac:
            return v[0] if single else v
        J = outJ.view(cdt) if cplx else outJ
        J = np.swapaxes(J.reshape(npts, N, K or 1), 1, 2)  # [p, k, d]
        return (v[0], J[0]) if single else (v, J)

    def _eval_torch(self, x, jac: bool, grad: bool):
        import torch

        L = _capi.lib()
        N, K = self.ndim, self.K
        cplx = self.is_complex
        E = (K or 1) * (2 if cplx else 1)
        if not x.is_cuda:
            raise ArgumentError("torch points must live on the GPU (pass numpy arrays for host batches)")
        if x.device.index != self.device:
            raise ArgumentError(f"points are on cuda:{x.device.index}, the polynomial on cuda:{self.device}")
        rdt = self._compute_dtype(np.float32 if x.dtype == torch.float32 else np.float64)
        tdt = torch.float32 if rdt == np.float32 else torch.float64
        pts = x.to(tdt).reshape(-1, N).contiguous()
        npts = pts.shape[0]
        h = self._handle(rdt)
        out = torch.empty((npts, E), dtype=tdt, device=x.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        if jac:
            if grad and K not in (None, 1):
                raise DimensionMismatch("chebgradient needs a scalar-valued polynomial")
            outJ = torch.empty((npts, N, E), dtype=tdt, device=x.device)
            rc = L.fastcheb_eval_jac(h, pts.data_ptr(), npts, out.data_ptr(), outJ.data_ptr(), _capi.MEM_DEVICE, stream)
        else:
            rc = L.fastcheb_eval(h, pts.data_ptr(), npts, out.data_ptr(), _capi.MEM_DEVICE, stream)
        if rc == _capi.EDOMAIN:
            bad = _capi.error_index()
            err = ArgumentError(f"{pts[bad].tolist()} not in domain {self.lb} to {self.ub}")
            err.index = bad
            raise err
        _check(rc, "fastcheb_eval: ")
        v = torch.view_as_complex(out.reshape(npts, -1, 2)) if cplx else out
        v = v.reshape(npts) if K is None else v.reshape(npts, K)
        if not jac:
            return v
        J = torch.view_as_complex(outJ.reshape(npts, N, -1, 2)) if cplx else outJ
        return v, J.reshape(npts, N, K or 1).transpose(1, 2)

    def __call__(self, x):  # eval.jl:63-70
        return self._eval(x, False)


def chebjacobian(c: ChebPoly, x):
    """(value, Jacobian) — eval.jl:147-155.  Jacobian shape (K, N) per point ((1, N) for scalar-valued c)."""
    return c._eval(x, True)


def chebgradient(c: ChebPoly, x):
    """(value, gradient) — eval.jl:168-177; DimensionMismatch for vector-valued c (eval.jl:170)."""
    if c.K is not None and c.K != 1:
        # go through the C ABI so the status code path is exercised; it raises DimensionMismatch
        pass
    v, J = c._eval(x, True, grad=True)
    if c.K is not None:
        v = v[..., 0]
    g = J[..., 0, :]
    if c.ndim == 1 and not _is_torch(x) and np.ndim(x) == 0:
        return v, g[..., 0]  # eval.jl:174-177
    return v, g


# ------------------------------------------------------------------------------------------- fit
def chebpoints(order, lb, ub, dtype=np.float64):
    """interp.jl:4-36.  Shape (n1..nN, N) for tuple orders, (n,) for a scalar order.  Index 0 along a
    dimension is the *upper* bound; order 0 gives the midpoint."""
    scalar = np.isscalar(order)
    order = (int(order),) if scalar else tuple(int(o) for o in order)
    lbv = np.atleast_1d(np.asarray(lb, dtype=dtype))
    ubv = np.atleast_1d(np.asarray(ub, dtype=dtype))
    if any(o < 0 for o in order):
        raise ArgumentError(f"invalid negative order {order}")  # interp.jl:10
    if not (len(order) == len(lbv) == len(ubv)):
        raise DimensionMismatch("order, lb and ub must have the same length")  # interp.jl:30
    dt = np.dtype(dtype).type
    axes = []
    for d, o in enumerate(order):
        i = np.array([1], dtype=dt) if o == 0 else np.arange(0, o + 1, dtype=dt)
        den = dt(2 if o == 0 else o)
        cth = np.cos(i * dt(np.pi) / den).astype(dt)
        axes.append((lbv[d] + (dt(1) + cth) * (ubv[d] - lbv[d]) * dt(0.5)).astype(dt))  # interp.jl:6
    if scalar:
        return axes[0]
    return np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1)


def _bounds(lb, ub):
    lbv, ubv = np.atleast_1d(np.asarray(lb)), np.atleast_1d(np.asarray(ub))
    if lbv.shape != ubv.shape or lbv.ndim != 1:
        raise DimensionMismatch("lb and ub must be vectors of the same length")
    return lbv, ubv


def chebcoefs(vals, ndim: int | None = None, device: int = 0, howmany: int | None = None, return_maxabs: bool = False):
    """chebcoefs(vals) — interp.jl:39-71 (no droptol).  `howmany`: vals has a leading batch axis of independent
    fits (the reference's way to spell this is a loop).  return_maxabs: also return max_k infnorm(C_k) of every
    fit (interp.jl:74-78), the quantity droptol scales its tolerance by."""
    vals = _float_vals(vals)
    batched = howmany is not None
    body = vals[0] if batched else vals
    if ndim is None:
        ndim = body.ndim
    K = body.shape[-1] if body.ndim == ndim + 1 else None
    cplx = np.iscomplexobj(vals)
    dims = body.shape[:ndim]
    rdt = _real_dtype(vals.dtype)
    if batched:
        nfit = vals.shape[0]
        if ndim == 1:  # (howmany, n[, K]) C-contiguous is already the Julia order of back-to-back 1-D fits
            flat = np.ascontiguousarray(vals).reshape(-1)
            flat = flat.view(rdt) if cplx else flat
        else:
            flat = np.concatenate([_to_julia_flat(v, ndim) for v in vals])
    else:
        flat = _to_julia_flat(vals, ndim)
        nfit = 1
    out = np.empty_like(flat)
    mx = np.zeros(nfit, dtype=np.float64) if return_maxabs else None
    rc = _capi.lib().fastcheb_fit(ndim, _capi.i64_array(dims), _dtype_code(rdt), int(cplx), K or 1, nfit,
                                  flat.ctypes.data, out.ctypes.data, mx.ctypes.data if return_maxabs else None,
                                  _capi.MEM_HOST, device, None)
    _check(rc, "fastcheb_fit: ")
    if batched:
        if ndim == 1:
            res = (out.view(vals.dtype) if cplx else out).reshape(vals.shape)
        else:
            per = out.size // nfit
            res = np.stack([_from_julia_flat(out[i * per:(i + 1) * per], dims, K, cplx) for i in range(nfit)])
    else:
        res = _from_julia_flat(out, dims, K, cplx)
    return (res, mx) if return_maxabs else res


def droptol(coefs, tol: float, ndim: int | None = None, device: int = 0):
    """droptol(coefs, tol) — interp.jl:77-93."""
    coefs = _float_vals(coefs)
    if ndim is None:
        ndim = coefs.ndim
    K = coefs.shape[-1] if coefs.ndim == ndim + 1 else None
    cplx = np.iscomplexobj(coefs)
    dims = coefs.shape[:ndim]
    flat = _to_julia_flat(coefs, ndim)
    newdims = (ctypes.c_int64 * ndim)()
    rc = _capi.lib().fastcheb_droptol_shape(ndim, _capi.i64_array(dims), _dtype_code(_real_dtype(coefs.dtype)), int(cplx),
                                            K or 1, flat.ctypes.data, float(tol), newdims, _capi.MEM_HOST, device, None)
    _check(rc, "fastcheb_droptol_shape: ")
    nd = tuple(int(v) for v in newdims)
    if nd == tuple(dims):
        return coefs
    return np.ascontiguousarray(coefs[tuple(slice(0, n) for n in nd)])  # interp.jl:92


def chebinterp(vals, lb, ub, tol: float | None = None, device: int = 0) -> ChebPoly:
    """chebinterp(vals, lb, ub; tol=eps) — interp.jl:95-99,140-141.  Fit, trim and handle creation all
    happen on the device (fastcheb_interp); the trimmed coefficients are copied back for `.coefs`."""
    lbv, ubv = _bounds(lb, ub)
    N = len(lbv)
    vals = _float_vals(vals)
    if vals.ndim not in (N, N + 1):
        raise DimensionMismatch("vals rank does not match the bounds")
    K = vals.shape[-1] if vals.ndim == N + 1 else None
    cplx = np.iscomplexobj(vals)
    rdt = _real_dtype(vals.dtype)
    flat = _to_julia_flat(vals, N)
    L = _capi.lib()
    h = ctypes.c_void_p()
    rc = L.fastcheb_interp(ctypes.byref(h), N, _capi.i64_array(vals.shape[:N]), _dtype_code(rdt), int(cplx), K or 1,
                           flat.ctypes.data, _capi.MEM_HOST, _capi.dbl_array(lbv), _capi.dbl_array(ubv),
                           -1.0 if tol is None else float(tol), device, None)
    _check(rc, "fastcheb_interp: ")
    newdims = (ctypes.c_int64 * N)()
    _check(L.fastcheb_dims(h, newdims))
    nd = tuple(int(v) for v in newdims)
    E = (K or 1) * (2 if cplx else 1)
    out = np.empty(int(np.prod(nd)) * E, dtype=rdt)
    _check(L.fastcheb_get_coefs(h, out.ctypes.data, _capi.MEM_HOST))
    coefs = _from_julia_flat(out, nd, K, cplx)
    return ChebPoly(coefs, lbv, ubv, device=device, _handle=h, _handle_dtype=rdt)


def chebinterp_v1(vals, lb, ub, tol: float | None = None, device: int = 0) -> ChebPoly:
    """interp.jl:155-156: the *first* axis of `vals` is the component axis."""
    return chebinterp(np.moveaxis(np.asarray(vals), 0, -1), lb, ub, tol=tol, device=device)

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
  chebinterp(f, order, lb, ub)  interp.jl:163-170 (host closure f on the host grid; a ChebPoly f stays on the device)
  chebinterp(f, lb, ub)         interp.jl:174-203 (adaptive order doubling)
  rrule, frule                  chainrules.jl:4-39 (batched pullback real(J' dy) / pushforward J dx with the
                                Jacobian contracted inside the kernel, never written to HBM)
  chebregression, chebvandermonde  regression.jl:5-73 (the Vandermonde matrix is assembled on the device in
                                one O(m P) pass; the dense QR solve `A \\ Y` stays a library call, LAPACK or cuSOLVER)
  roots, colleague_matrix       roots.jl:55-59, 88-110 (re-interpolation on the device, <= 50x50
                                eigenproblems on the host, as the reference leaves them to LAPACK)

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
from ._capi import ALGO_AUTO, ALGO_CLENSHAW, ALGO_DMMA, ALGO_DMMA_RING, ALGO_PRODUCT  # noqa: F401

__all__ = ["MultiChebPoly", "chebcoefs_multi", "rrule", "frule", "chebregression", "chebvandermonde", "eval_many", "ChebPoly", "chebinterp", "chebinterp_v1", "chebjacobian", "chebgradient", "chebpoints", "chebcoefs",
           "droptol", "roots", "colleague_matrix", "ArgumentError", "DomainError", "DimensionMismatch", "CudaError"]

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

    def product_check(self, rdt=np.float64):
        """Self-check of the 1-D product-basis path against the Clenshaw kernel (fastcheb_product_check): worst value
        difference over 16 eps sum|c| and worst derivative difference over 16 eps sum|c_k| k^2 2/(ub-lb); AUTO uses the
        path when they are <= 0.25."""
        dv, dj = ctypes.c_double(0), ctypes.c_double(0)
        _check(_capi.lib().fastcheb_product_check(self._handle(rdt), ctypes.byref(dv), ctypes.byref(dj)), "fastcheb_product_check: ")
        return dv.value, dj.value

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


class MultiChebPoly:
    """A ChebPoly replicated on several GPUs of the node and evaluated from ONE host process (fastcheb_multi_*):
    `c.(ys)`, chebjacobian and chebgradient block-partition the points over the devices, results come back in one
    array in point order, the ArgumentError names the first offending point of the whole batch (eval.jl:64).
    numpy points -> host path (every device stages its block); torch CUDA points on devices[0] -> the other devices
    read their block from, and write their results into, devices[0]'s memory over NVLink."""

    def __init__(self, coefs, lb, ub, devices=None):
        L = _capi.lib()
        self.base = ChebPoly(coefs, lb, ub, device=(devices[0] if devices else 0))
        ndev = L.fastcheb_device_count()
        self.devices = list(range(ndev)) if devices is None else [int(d) for d in devices]
        self._multi = {}

    def _handle(self, rdt):
        rdt = np.dtype(rdt)
        h = self._multi.get(rdt)
        if h is None:
            b = self.base
            flat = np.ascontiguousarray(_to_julia_flat(b.coefs, b.ndim).astype(rdt, copy=False))
            h = ctypes.c_void_p()
            devs = (ctypes.c_int * len(self.devices))(*self.devices)
            rc = _capi.lib().fastcheb_multi_create(ctypes.byref(h), len(self.devices), devs, b.ndim, _capi.i64_array(b.dims),
                                                   _dtype_code(rdt), int(b.is_complex), b.K or 1, flat.ctypes.data,
                                                   _capi.MEM_HOST, _capi.dbl_array(b.lb), _capi.dbl_array(b.ub))
            _check(rc, "fastcheb_multi_create: ")
            self._multi[rdt] = h
        return h

    def close(self):
        for h in self._multi.values():
            _capi.lib().fastcheb_multi_destroy(h)
        self._multi = {}
        self.base.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _eval(self, x, jac: bool, grad: bool = False):
        L = _capi.lib()
        b = self.base
        N, K, cplx = b.ndim, b.K, b.is_complex
        E = (K or 1) * (2 if cplx else 1)
        if _is_torch(x):
            import torch

            if not x.is_cuda or x.device.index != self.devices[0]:
                raise ArgumentError(f"device-resident points must live on cuda:{self.devices[0]} (devices[0])")
            rdt = b._compute_dtype(np.float32 if x.dtype == torch.float32 else np.float64)
            tdt = torch.float32 if rdt == np.float32 else torch.float64
            pts = x.to(tdt).reshape(-1, N).contiguous()
            npts = pts.shape[0]
            out = torch.empty((npts, E), dtype=tdt, device=x.device)
            outJ = torch.empty((npts, N, E), dtype=tdt, device=x.device) if jac else None
            torch.cuda.current_stream(x.device).synchronize()  # the workers use their own streams
            p_in, p_v, p_J, mem = pts.data_ptr(), out.data_ptr(), (outJ.data_ptr() if jac else None), _capi.MEM_DEVICE
        else:
            xa = np.asarray(x)
            rdt = b._compute_dtype(xa.dtype if xa.dtype.kind == "f" else np.float64)
            pts = np.ascontiguousarray(xa.astype(rdt, copy=False).reshape(-1, N))
            npts = pts.shape[0]
            out = np.empty((npts, E), dtype=rdt)
            outJ = np.empty((npts, N, E), dtype=rdt) if jac else None
            p_in, p_v, p_J, mem = pts.ctypes.data, out.ctypes.data, (outJ.ctypes.data if jac else None), _capi.MEM_HOST
        h = self._handle(rdt)
        if jac:
            fn = L.fastcheb_multi_eval_grad if grad else L.fastcheb_multi_eval_jac
            rc = fn(h, p_in, npts, p_v, p_J, mem)
        else:
            rc = L.fastcheb_multi_eval(h, p_in, npts, p_v, mem)
        if rc == _capi.EDOMAIN:
            bad = _capi.error_index()
            err = ArgumentError(f"{pts[bad].tolist()} not in domain {b.lb} to {b.ub}")  # eval.jl:64
            err.index = bad
            raise err
        _check(rc, "fastcheb_multi_eval: ")
        if _is_torch(x):
            import torch

            v = torch.view_as_complex(out.reshape(npts, -1, 2)) if cplx else out
            v = v.reshape(npts) if K is None else v.reshape(npts, K)
            if not jac:
                return v
            J = torch.view_as_complex(outJ.reshape(npts, N, -1, 2)) if cplx else outJ
            return v, J.reshape(npts, N, K or 1).transpose(1, 2)
        cdt = np.complex64 if rdt == np.float32 else np.complex128
        v = out.view(cdt) if cplx else out
        v = v.reshape(npts) if K is None else v.reshape(npts, K)
        if not jac:
            return v
        J = outJ.view(cdt) if cplx else outJ
        return v, np.swapaxes(J.reshape(npts, N, K or 1), 1, 2)

    def __call__(self, x):
        return self._eval(x, False)

    # chebjacobian / chebgradient accept this type through the same functions
    @property
    def ndim(self):
        return self.base.ndim

    @property
    def K(self):
        return self.base.K

    @property
    def lb(self):
        return self.base.lb

    @property
    def ub(self):
        return self.base.ub


def chebcoefs_multi(vals, devices=None, return_maxabs: bool = False):
    """A batch of independent 1-D fits (leading batch axis) split over the GPUs of the node from one process
    (fastcheb_multi_fit; SURVEY.md 8e: fits are sharded per GPU, a single fit is never split)."""
    vals = _float_vals(vals)
    if vals.ndim not in (2, 3):
        raise DimensionMismatch("vals must be (howmany, n) or (howmany, n, K)")
    L = _capi.lib()
    devs = list(range(L.fastcheb_device_count())) if devices is None else [int(d) for d in devices]
    howmany, n = vals.shape[:2]
    K = vals.shape[2] if vals.ndim == 3 else None
    cplx = np.iscomplexobj(vals)
    rdt = _real_dtype(vals.dtype)
    flat = np.ascontiguousarray(vals).reshape(-1)
    flat = flat.view(rdt) if cplx else flat
    out = np.empty_like(flat)
    mx = np.zeros(howmany, dtype=np.float64) if return_maxabs else None
    rc = L.fastcheb_multi_fit(len(devs), (ctypes.c_int * len(devs))(*devs), 1, _capi.i64_array((n,)), _dtype_code(rdt), int(cplx),
                              K or 1, howmany, flat.ctypes.data, out.ctypes.data, mx.ctypes.data if return_maxabs else None,
                              _capi.MEM_HOST)
    _check(rc, "fastcheb_multi_fit: ")
    res = (out.view(vals.dtype) if cplx else out).reshape(vals.shape)
    return (res, mx) if return_maxabs else res


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
    if c.ndim == 1 and (x.dim() if _is_torch(x) else np.ndim(x)) <= 1:
        return v, g[..., 0]  # eval.jl:174-177: a scalar derivative per scalar point (a 1-D array is a batch of them)
    return v, g


def eval_many(coefs, lb, ub, pts, deriv: bool = False, device: int = 0):
    """Many independent 1-D polynomials, each at its own points, in one launch: `[c.(xs) for (c, xs) in zip(cs, xss)]`
    (eval.jl:63-70 per point; with deriv=True also chebgradient's derivative, eval.jl:174-177).

    coefs: (howmany, n) or (howmany, n, K) real/complex — the layout chebcoefs(..., howmany=) returns;
    lb, ub: scalars (shared box) or length-howmany arrays; pts: (howmany, M) reals, or (M,) reals SHARED by all
    polynomials (with a shared box: the batch is then a matrix product on the FP64 tensor cores,
    fastcheb_eval_many_shared; n <= 65).
    Returns values (howmany, M[, K]) [and derivatives of the same shape]."""
    coefs = _float_vals(coefs)
    if coefs.ndim not in (2, 3):
        raise DimensionMismatch("coefs must be (howmany, n) or (howmany, n, K)")
    howmany, n = coefs.shape[:2]
    K = coefs.shape[2] if coefs.ndim == 3 else None
    cplx = np.iscomplexobj(coefs)
    rdt = _real_dtype(coefs.dtype)
    pts = np.ascontiguousarray(np.asarray(pts), dtype=rdt)
    lbv, ubv = np.atleast_1d(np.asarray(lb, dtype=np.float64)), np.atleast_1d(np.asarray(ub, dtype=np.float64))
    per = int(lbv.size > 1 or ubv.size > 1)
    if pts.ndim == 1 and not per and n <= 65:
        return _eval_many_shared(coefs, lbv, ubv, pts, deriv, device, howmany, n, K, cplx, rdt)
    if pts.ndim == 1:
        pts = np.ascontiguousarray(np.broadcast_to(pts, (howmany, pts.shape[0])))
    if pts.ndim != 2 or pts.shape[0] != howmany:
        raise DimensionMismatch("pts must be (howmany, M) or (M,)")
    M = pts.shape[1]
    if per:
        lbv, ubv = np.broadcast_to(lbv, (howmany,)).copy(), np.broadcast_to(ubv, (howmany,)).copy()
    flat = np.ascontiguousarray(coefs).reshape(-1)
    flat = flat.view(rdt) if cplx else flat
    E = (K or 1) * (2 if cplx else 1)
    out = np.empty((howmany, M, E), dtype=rdt)
    outd = np.empty((howmany, M, E), dtype=rdt) if deriv else None
    rc = _capi.lib().fastcheb_eval_many(howmany, n, _dtype_code(rdt), int(cplx), K or 1, flat.ctypes.data,
                                        lbv.ctypes.data, ubv.ctypes.data, per, pts.ctypes.data, M, out.ctypes.data,
                                        outd.ctypes.data if deriv else None, None, _capi.MEM_HOST, device, None)
    if rc == _capi.EDOMAIN:
        bad = _capi.error_index()
        f, i = divmod(bad, M)
        err = ArgumentError(f"{pts[f, i]} not in domain {lbv[f if per else 0]} to {ubv[f if per else 0]} (polynomial {f})")
        err.index = bad
        raise err
    _check(rc, "fastcheb_eval_many: ")

    def shape(a):
        a = a.view(np.complex64 if rdt == np.float32 else np.complex128) if cplx else a
        return a.reshape(howmany, M) if K is None else a.reshape(howmany, M, K)

    return (shape(out), shape(outd)) if deriv else shape(out)


def _eval_many_shared(coefs, lbv, ubv, pts, deriv, device, howmany, n, K, cplx, rdt):
    M = pts.shape[0]
    flat = np.ascontiguousarray(coefs).reshape(-1)
    flat = flat.view(rdt) if cplx else flat
    E = (K or 1) * (2 if cplx else 1)
    out = np.empty((howmany, M, E), dtype=rdt)
    outd = np.empty((howmany, M, E), dtype=rdt) if deriv else None
    rc = _capi.lib().fastcheb_eval_many_shared(howmany, n, _dtype_code(rdt), int(cplx), K or 1, flat.ctypes.data, lbv.ctypes.data,
                                               ubv.ctypes.data, pts.ctypes.data, M, out.ctypes.data,
                                               outd.ctypes.data if deriv else None, None, _capi.MEM_HOST, device, None)
    if rc == _capi.EDOMAIN:
        bad = _capi.error_index()
        err = ArgumentError(f"{pts[bad]} not in domain {lbv[0]} to {ubv[0]}")
        err.index = bad
        raise err
    _check(rc, "fastcheb_eval_many_shared: ")

    def shape(a):
        a = a.view(np.complex64 if rdt == np.float32 else np.complex128) if cplx else a
        return a.reshape(howmany, M) if K is None else a.reshape(howmany, M, K)

    return (shape(out), shape(outd)) if deriv else shape(out)


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


def _handle_to_poly(h, lbv, ubv, rdt, K, cplx, device) -> ChebPoly:
    L = _capi.lib()
    N = len(lbv)
    newdims = (ctypes.c_int64 * N)()
    _check(L.fastcheb_dims(h, newdims))
    nd = tuple(int(v) for v in newdims)
    E = (K or 1) * (2 if cplx else 1)
    out = np.empty(int(np.prod(nd)) * E, dtype=rdt)
    _check(L.fastcheb_get_coefs(h, out.ctypes.data, _capi.MEM_HOST))
    coefs = _from_julia_flat(out, nd, K, cplx)
    return ChebPoly(coefs, lbv, ubv, device=device, _handle=h, _handle_dtype=rdt)


def _eps_bounds(lbv, ubv):
    """epsbounds (interp.jl:158-160): eps of the float type of the bounds."""
    td = np.result_type(np.asarray(lbv).dtype, np.asarray(ubv).dtype)
    return float(np.finfo(td if td.kind == "f" else np.float64).eps)


def chebinterp(vals, *args, tol: float | None = None, device: int = 0, min_order=None, max_order=None) -> ChebPoly:
    """chebinterp(vals, lb, ub; tol=eps)         — interp.jl:95-99,140-141: fit, trim and handle creation all happen
                                                   on the device (fastcheb_interp).
    chebinterp(f, order, lb, ub; tol=eps)        — interp.jl:163-170.  A host callable `f` is evaluated on the host grid
                                                   (as in the reference); if `f` is a ChebPoly the grid, its evaluation
                                                   and the fit stay on the device (fastcheb_reinterp).
    chebinterp(f, lb, ub; tol=5eps, min_order, max_order) — interp.jl:174-203: order doubling until every dimension's
                                                   fitted order is below the trial order."""
    if callable(vals):
        f = vals
        if len(args) == 3:
            return _chebinterp_f(f, args[0], args[1], args[2], tol, device)
        if len(args) == 2:
            return _chebinterp_adaptive(f, args[0], args[1], tol, device, min_order, max_order)
        raise ArgumentError("chebinterp(f, [order,] lb, ub)")
    if len(args) != 2:
        raise ArgumentError("chebinterp(vals, lb, ub)")
    lb, ub = args
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
    return _handle_to_poly(h, lbv, ubv, rdt, K, cplx, device)


def _chebinterp_f(f, order, lb, ub, tol, device) -> ChebPoly:
    scalar = np.isscalar(order)
    lbv, ubv = _bounds(lb, ub)
    orders = (int(order),) * 1 if scalar else tuple(int(o) for o in order)
    if any(o < 0 for o in orders):
        raise ArgumentError(f"invalid negative order {orders}")  # interp.jl:10
    if len(orders) != len(lbv):
        raise DimensionMismatch("order, lb and ub must have the same length")
    if tol is None:
        tol = _eps_bounds(lbv, ubv)  # interp.jl:163: tol = epsbounds(lb, ub)
    if isinstance(f, ChebPoly):
        if f.ndim != len(lbv):
            raise DimensionMismatch("the polynomial and the new box differ in dimension")
        rdt = _real_dtype(f.coefs.dtype)
        h = ctypes.c_void_p()
        rc = _capi.lib().fastcheb_reinterp(ctypes.byref(h), f._handle(rdt), _capi.i64_array(orders), _capi.dbl_array(lbv),
                                           _capi.dbl_array(ubv), float(tol), None)
        _check(rc, "fastcheb_reinterp: ")
        return _handle_to_poly(h, lbv, ubv, rdt, f.K, f.is_complex, f.device)
    fdt = np.result_type(lbv.dtype, ubv.dtype)
    fdt = fdt if fdt.kind == "f" else np.float64
    x = chebpoints(orders[0] if scalar else orders, lbv[0] if scalar else lbv, ubv[0] if scalar else ubv, dtype=fdt)
    if scalar:
        vals = np.asarray([f(xi) for xi in x])
    else:
        flat = x.reshape(-1, len(orders))
        vals = np.asarray([f(xi) for xi in flat])
        vals = vals.reshape(x.shape[:-1] + vals.shape[1:])
    return chebinterp(vals, lbv, ubv, tol=tol, device=device)


def _chebinterp_adaptive(f, lb, ub, tol, device, min_order, max_order) -> ChebPoly:
    lbv, ubv = _bounds(lb, ub)
    N = len(lbv)
    if tol is None:
        tol = 5 * _eps_bounds(lbv, ubv)
    if not tol > 0:
        raise ArgumentError(f"tolerance {tol} must be > 0")  # interp.jl:178
    big = np.iinfo(np.int64).max
    mn = np.full(N, 8, dtype=np.int64) if min_order is None else np.atleast_1d(np.asarray(min_order, dtype=np.int64))
    mx = np.full(N, big, dtype=np.int64) if max_order is None else np.atleast_1d(np.asarray(max_order, dtype=np.int64))
    if not (len(mn) == len(mx) == N):
        raise DimensionMismatch(f"dimensions must all == {N}")
    if not np.all(mn > 0):
        raise ArgumentError(f"minimum order {mn} must be > 0")
    if not np.all(mx >= 0):
        raise ArgumentError(f"maximum order {mx} must be >= 0")
    order = np.minimum(mn, mx)
    scalar = np.ndim(lb) == 0
    while True:
        c = _chebinterp_f(f, int(order[0]) if scalar else tuple(int(o) for o in order), lb, ub, tol, device)
        size = np.asarray(c.dims)
        done = (size - 1 < order) | (order == mx)                      # interp.jl:189
        if np.all(done):
            return c
        nxt = np.array([1 << int(2 * o - 1).bit_length() if 2 * o > 1 else 1 for o in order], dtype=np.int64)  # nextpow(2, 2o)
        order = np.where(done, order, np.minimum(mx, nxt))


def chebinterp_v1(vals, lb, ub, tol: float | None = None, device: int = 0) -> ChebPoly:
    """interp.jl:155-156: the *first* axis of `vals` is the component axis."""
    return chebinterp(np.moveaxis(np.asarray(vals), 0, -1), lb, ub, tol=tol, device=device)



# ------------------------------------------------------------------------------------------- roots
def _colleague(coefs: np.ndarray) -> np.ndarray:
    """Colleague matrix on [-1, 1] of a coefficient vector whose last entry is non-zero (roots.jl:13-33): the
    upper-Hessenberg matrix with halves on both off-diagonals, a one at (2,1), and the last column corrected by
    -c[:-1] / (2 c[-1])."""
    n = len(coefs)
    if n <= 1:
        return np.zeros((0, 0))
    if coefs[-1] == 0:
        raise ArgumentError("trailing coefficient must be nonzero")
    if n == 2:
        return np.array([[-coefs[0] / coefs[1]]], dtype=float)
    m = n - 1
    C = np.zeros((m, m))
    i = np.arange(m - 1)
    C[i, i + 1] = 0.5
    C[i + 1, i] = 0.5
    C[1, 0] = 1.0
    C[:, -1] -= coefs[:-1] / (2 * coefs[-1])
    return C


def colleague_matrix(c: ChebPoly, tol: float | None = None) -> np.ndarray:
    """colleague_matrix(c; tol=5eps) — roots.jl:55-59: eigenvalues are the roots of the 1-D polynomial on [lb, ub];
    trailing coefficients below tol * sum|c| are dropped."""
    if c.ndim != 1 or c.K is not None:
        raise ArgumentError("colleague_matrix needs a scalar-valued 1-D ChebPoly")
    co = np.asarray(c.coefs)
    if tol is None:
        tol = 5 * float(np.finfo(_real_dtype(co.dtype)).eps)
    abstol = np.abs(co).sum() * tol
    big = np.nonzero(np.abs(co) >= abstol)[0]
    n = int(big[-1]) + 1 if len(big) else 1
    C = _colleague(co[:n].astype(float))
    lb, ub = float(c.lb[0]), float(c.ub[0])
    C = C * ((ub - lb) / 2)
    C[np.diag_indices_from(C)] += (ub + lb) / 2
    return C


def roots(c: ChebPoly, tol: float | None = None, maxsize: int = 50) -> np.ndarray:
    """roots(c; tol=5eps, maxsize=50) — roots.jl:88-110: sorted real roots of a real 1-D ChebPoly on [lb, ub].
    Small problems are colleague-matrix eigenvalues (host LAPACK, like the reference); larger ones are split at the
    chebfun point and *re-interpolated on the device* (fastcheb_reinterp: grid -> batched evaluation -> DCT-I fit ->
    droptol), recursively, with the coefficients never leaving the GPU except for the <= maxsize eigenproblems."""
    if c.ndim != 1 or c.K is not None or c.is_complex:
        raise ArgumentError("roots needs a real scalar-valued 1-D ChebPoly")
    co = np.asarray(c.coefs)
    if tol is None:
        tol = 5 * float(np.finfo(_real_dtype(co.dtype)).eps)
    if not tol > 0:
        raise ArgumentError(f"tolerance {tol} for truncating coefficients must be > 0")
    if not maxsize > 0:
        raise ArgumentError(f"maxsize {maxsize} must be > 0")
    lb, ub = float(c.lb[0]), float(c.ub[0])
    abstol = np.abs(co).max() * tol
    big = np.nonzero(np.abs(co) >= abstol)[0]
    n = int(big[-1]) + 1 if len(big) else 1
    if n <= maxsize:
        C = _colleague(co[:n].astype(float))
        lam = np.linalg.eigvals(C) if C.size else np.zeros(0)
        lam = (lam + 1) * ((ub - lb) / 2) + lb
        htol = float(np.spacing(ub - lb)) * 100                         # roots.jl:63: eps(float(ub - lb)) * 100
        keep = (np.abs(lam.imag) < htol) & (lam.real >= lb - htol) & (lam.real <= ub + htol)
        return np.sort(np.clip(lam.real[keep], lb, ub))
    split = 1.004849834917525 * ((ub - lb) / 2) + lb                  # roots.jl:102 (chebfun's split point)
    order = 1 << int(n - 2).bit_length() if n > 2 else 1               # nextpow(2, n-1)
    c1 = chebinterp(c, order, lb, split, tol=2 * tol)
    c2 = chebinterp(c, order, split, ub, tol=2 * tol)
    return np.concatenate([roots(c1, tol=2 * tol, maxsize=maxsize), roots(c2, tol=2 * tol, maxsize=maxsize)])


# ------------------------------------------------------------------------------------------- regression
def chebvandermonde(x, lb, ub, order, device: int = 0):
    """chebvandermonde(x, lb, ub, order) — regression.jl:5-46: the m x prod(order+1) Chebyshev-Vandermonde matrix,
    columns in column-major multi-index order.  numpy in -> numpy (Fortran-ordered) out; a torch CUDA tensor in ->
    a torch tensor on the same device (column-major strides) out, without touching the host."""
    scalar = np.isscalar(order)
    orders = (int(order),) if scalar else tuple(int(o) for o in order)
    N = len(orders)
    if any(o < 0 for o in orders):
        raise ArgumentError(f"order {orders} must be nonnegative")  # regression.jl:30
    lbv, ubv = _bounds(lb, ub)
    if len(lbv) != N:
        raise DimensionMismatch("order, lb and ub must have the same length")
    dims = tuple(o + 1 for o in orders)
    P = int(np.prod(dims))
    L = _capi.lib()
    if _is_torch(x):
        import torch

        tdt = torch.float32 if x.dtype == torch.float32 else torch.float64
        pts = x.to(tdt).reshape(-1, N).contiguous()
        m = pts.shape[0]
        At = torch.empty((P, m), dtype=tdt, device=x.device)  # row i of At = column i of A
        rc = L.fastcheb_vandermonde(N, _capi.i64_array(dims), _capi.F32 if tdt == torch.float32 else _capi.F64,
                                    pts.data_ptr(), m, _capi.dbl_array(lbv), _capi.dbl_array(ubv), At.data_ptr(), m,
                                    _capi.MEM_DEVICE, x.device.index,
                                    ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
        A = At.t()
        ptsh = None
    else:
        xa = np.asarray(x)
        rdt = np.float32 if xa.dtype == np.float32 and lbv.dtype != np.float64 else np.float64
        ptsh = np.ascontiguousarray(xa.astype(rdt, copy=False).reshape(-1, N))
        m = ptsh.shape[0]
        A = np.empty((m, P), dtype=rdt, order="F")
        rc = L.fastcheb_vandermonde(N, _capi.i64_array(dims), _dtype_code(rdt), ptsh.ctypes.data, m,
                                    _capi.dbl_array(lbv), _capi.dbl_array(ubv), A.ctypes.data, m, _capi.MEM_HOST, device, None)
    if rc == _capi.EDOMAIN:
        bad = _capi.error_index()
        err = ArgumentError(f"{ptsh[bad] if ptsh is not None else 'point %d' % bad} not in domain [{lbv},{ubv}]")  # regression.jl:33
        err.index = bad
        raise err
    _check(rc, "fastcheb_vandermonde: ")
    return A


def chebregression(x, y, *args, device: int = 0) -> ChebPoly:
    """chebregression(x, y, [lb, ub,] order) — regression.jl:52-107: least-squares fit of a Chebyshev polynomial of the
    given order to scattered data.  x: (m,) reals (1-D) or (m, N); y: (m,) or (m, K) (vector-valued fit); lb/ub default
    to the coordinate-wise minimum/maximum of x.  The Vandermonde matrix is assembled on the GPU; `A \\ Y` is a dense QR
    least-squares solve (LAPACK on the host for numpy input — as in the reference — or cuSOLVER through torch for
    device-resident input)."""
    if len(args) == 3:
        lb, ub, order = args
    elif len(args) == 1:
        order = args[0]
        lb = ub = None
    else:
        raise ArgumentError("chebregression(x, y, [lb, ub,] order)")
    scalar = np.isscalar(order)
    orders = (int(order),) if scalar else tuple(int(o) for o in order)
    N = len(orders)
    on_dev = _is_torch(x)
    xa = x if on_dev else np.asarray(x)
    xm = xa.reshape(-1, N)
    m = xm.shape[0]
    if lb is None:  # regression.jl:91-96: bounds from the data
        lb = xm.min(0)[0].cpu().numpy() if on_dev else xm.min(axis=0)
        ub = xm.max(0)[0].cpu().numpy() if on_dev else xm.max(axis=0)
    lbv, ubv = _bounds(lb, ub)
    ya = y if _is_torch(y) else np.asarray(y)
    if ya.shape[0] != m:
        raise DimensionMismatch("x and y differ in length")  # regression.jl:54
    P = int(np.prod([o + 1 for o in orders]))
    if m < P:
        raise ArgumentError(f"not enough data points {m} to fit to order {orders}")  # regression.jl:55
    K = ya.shape[1] if ya.ndim == 2 else None
    A = chebvandermonde(xa, lbv, ubv, orders, device=device)
    if on_dev:
        import torch

        Y = (ya if _is_torch(ya) else torch.as_tensor(np.asarray(ya), device=x.device)).to(A.dtype).reshape(m, -1)
        C = torch.linalg.lstsq(A, Y).solution.cpu().numpy()          # cuSOLVER QR
    else:
        Y = np.asarray(ya).reshape(m, -1)
        cdt = np.result_type(A.dtype, Y.dtype)
        C = np.linalg.lstsq(A.astype(cdt, copy=False), Y.astype(cdt, copy=False), rcond=None)[0]   # LAPACK, like `A \\ Y` (:67)
    dims = tuple(o + 1 for o in orders)
    coefs = C.reshape(dims + (C.shape[1],), order="F") if K is not None else C[:, 0].reshape(dims, order="F")
    return ChebPoly(coefs, lbv, ubv, device=device if not on_dev else x.device.index)


# ------------------------------------------------------------------------------------------- AD rules
def _tangent_call(c: ChebPoly, x, tan, mode: int):
    """(values, contraction) for a batch of points: mode 1 = pullback real(J' dy), mode 2 = pushforward J dx."""
    L = _capi.lib()
    N, K, cplx = c.ndim, c.K, c.is_complex
    E = (K or 1) * (2 if cplx else 1)
    xa = np.asarray(x)
    rdt = c._compute_dtype(xa.dtype if xa.dtype.kind == "f" else np.float64)
    single = (xa.ndim == 0) if N == 1 else (xa.ndim == 1)
    pts = np.ascontiguousarray(xa.astype(rdt, copy=False).reshape(-1, N))
    npts = pts.shape[0]
    ta = np.asarray(tan)
    if mode == 1:
        cdt = np.complex64 if rdt == np.float32 else np.complex128
        ta = np.ascontiguousarray(ta.astype(cdt if cplx else rdt).reshape(npts, K or 1))
        tflat = ta.view(rdt).reshape(npts, E)
    else:
        if np.iscomplexobj(ta):
            raise ArgumentError("the tangent of a real point must be real")
        tflat = np.ascontiguousarray(ta.astype(rdt).reshape(npts, N))
    h = c._handle(rdt)
    out_v = np.empty((npts, E), dtype=rdt)
    out_t = np.empty((npts, N if mode == 1 else E), dtype=rdt)
    fn = L.fastcheb_eval_vjp if mode == 1 else L.fastcheb_eval_jvp
    rc = fn(h, pts.ctypes.data, npts, tflat.ctypes.data, out_v.ctypes.data, out_t.ctypes.data, _capi.MEM_HOST, None)
    if rc == _capi.EDOMAIN:
        bad = _capi.error_index()
        err = ArgumentError(f"{pts[bad]} not in domain {c.lb} to {c.ub}")
        err.index = bad
        raise err
    _check(rc, "fastcheb_eval_vjp/jvp: ")
    cdt = np.complex64 if rdt == np.float32 else np.complex128
    v = out_v.view(cdt) if cplx else out_v
    v = v.reshape(npts) if K is None else v.reshape(npts, K)
    if mode == 2:
        t = out_t.view(cdt) if cplx else out_t
        t = t.reshape(npts) if K is None else t.reshape(npts, K)
    else:
        t = out_t[:, 0] if (N == 1 and xa.ndim <= 1 and (xa.ndim == 0 or xa.shape[-1:] != (1,))) else out_t
    return (v[0], t[0]) if single else (v, t)


def rrule(c: ChebPoly, x):
    """ChainRulesCore.rrule(c, x) — chainrules.jl:4-16, batched: returns (y, pullback) with
    pullback(dy) = real(J' dy) per point (shape of x).  Each pullback call is one fused kernel launch
    (fastcheb_eval_vjp); there is no rule for perturbing the ChebPoly itself (as in the reference)."""
    y = c(x)

    def chebpoly_pullback(dy):
        return _tangent_call(c, x, dy, 1)[1]

    return y, chebpoly_pullback


def frule(dx, c: ChebPoly, x, dself: ChebPoly | None = None, dlb=None, dub=None):
    """ChainRulesCore.frule((Δself, Δx), c, x) — chainrules.jl:18-39, batched: returns (y, ẏ) with ẏ = J Δx per point.
    `dself` (a ChebPoly holding Δcoefs on c's box), `dlb`, `dub` are the optional tangents of the polynomial itself:
    Δx' = Δx + (d2 - 1) Δlb - d2 Δub with d2 = (x - lb)/(ub - lb) (:31-32) and ẏ += Δcoefs(x) (:35-37)."""
    xa = np.asarray(x)
    if xa.dtype.kind != "f":
        xa = xa.astype(np.float64)
    fdt = xa.dtype  # Float32 points keep the Float32 promotion path (eval.jl:25)
    dxa = np.broadcast_to(np.asarray(dx, dtype=fdt), xa.shape).copy()
    if dlb is not None or dub is not None:
        lb, ub = c.lb.astype(fdt), c.ub.astype(fdt)
        d2 = (xa - lb) / (ub - lb)
        dxa = dxa + (d2 - 1) * (0.0 if dlb is None else np.asarray(dlb, dtype=fdt)) - d2 * (0.0 if dub is None else np.asarray(dub, dtype=fdt))
    y, ydot = _tangent_call(c, xa, dxa, 2)
    if dself is not None:
        ydot = ydot + dself(xa)
    return y, ydot

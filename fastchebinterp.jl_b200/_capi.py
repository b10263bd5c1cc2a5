"""ctypes binding of libfastcheb_cuda.so (include/fastcheb.h).

This is the Python stand-in for the `ccall` sites of the Julia wrapper (julia/src/cuda.jl): the same
entry points, the same argument order.  There is no fallback: if the shared library is missing the
import of the product package fails loudly.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfastcheb_cuda.so")

F32, F64 = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
OK, EDOMAIN, ENONFINITE, EDIMMISMATCH, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED = range(8)
ALGO_AUTO, ALGO_CLENSHAW, ALGO_DMMA, ALGO_DMMA_RING, ALGO_PRODUCT = 0, 1, 2, 3, 4
MAX_NDIM = 6

# every symbol include/fastcheb.h declares (tests check the library exports all of them)
SYMBOLS = [
    "fastcheb_version", "fastcheb_launch_count", "fastcheb_device_count", "fastcheb_last_error", "fastcheb_error_index",
    "fastcheb_create", "fastcheb_destroy", "fastcheb_ndim", "fastcheb_dims", "fastcheb_dtype",
    "fastcheb_is_complex", "fastcheb_ncomp", "fastcheb_device", "fastcheb_get_coefs", "fastcheb_set_algo",
    "fastcheb_get_algo", "fastcheb_product_check", "fastcheb_eval", "fastcheb_eval_jac", "fastcheb_eval_grad", "fastcheb_eval_jac_tuple",
    "fastcheb_eval_async", "fastcheb_eval_jac_async", "fastcheb_eval_many", "fastcheb_eval_many_shared", "fastcheb_eval_vjp", "fastcheb_eval_jvp", "fastcheb_fit", "fastcheb_fit_async", "fastcheb_droptol_shape", "fastcheb_interp", "fastcheb_reinterp", "fastcheb_vandermonde",
    "fastcheb_multi_create", "fastcheb_multi_destroy", "fastcheb_multi_ngpus", "fastcheb_multi_device", "fastcheb_multi_poly",
    "fastcheb_multi_eval", "fastcheb_multi_eval_jac", "fastcheb_multi_eval_grad", "fastcheb_multi_eval_jac_tuple", "fastcheb_multi_fit",
    "fastcheb_peer_alloc", "fastcheb_peer_free", "fastcheb_peer_open", "fastcheb_peer_close", "fastcheb_peer_read",
]

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C fastchebinterp.jl_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`).  There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, i32, dbl = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
    i64p, dblp = ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_double)
    sig = {
        "fastcheb_version": (i32, []),
        "fastcheb_device_count": (i32, []),
        "fastcheb_launch_count": (i64, []),
        "fastcheb_last_error": (ctypes.c_size_t, [ctypes.c_char_p, ctypes.c_size_t]),
        "fastcheb_error_index": (i32, [i64p]),
        "fastcheb_create": (i32, [ctypes.POINTER(vp), i32, i64p, i32, i32, i32, vp, i32, dblp, dblp, i32]),
        "fastcheb_destroy": (i32, [vp]),
        "fastcheb_ndim": (i32, [vp]),
        "fastcheb_dims": (i32, [vp, i64p]),
        "fastcheb_dtype": (i32, [vp]),
        "fastcheb_is_complex": (i32, [vp]),
        "fastcheb_ncomp": (i32, [vp]),
        "fastcheb_device": (i32, [vp]),
        "fastcheb_get_coefs": (i32, [vp, vp, i32]),
        "fastcheb_set_algo": (i32, [vp, i32]),
        "fastcheb_get_algo": (i32, [vp, i32]),
        "fastcheb_product_check": (i32, [vp, dblp, dblp]),
        "fastcheb_eval": (i32, [vp, vp, i64, vp, i32, vp]),
        "fastcheb_eval_jac": (i32, [vp, vp, i64, vp, vp, i32, vp]),
        "fastcheb_eval_grad": (i32, [vp, vp, i64, vp, vp, i32, vp]),
        "fastcheb_eval_jac_tuple": (i32, [vp, vp, i64, vp, i32, vp]),
        "fastcheb_eval_async": (i32, [vp, vp, i64, vp, vp, vp]),
        "fastcheb_eval_jac_async": (i32, [vp, vp, i64, vp, vp, vp, vp]),
        "fastcheb_fit": (i32, [i32, i64p, i32, i32, i32, i64, vp, vp, vp, i32, i32, vp]),
        "fastcheb_reinterp": (i32, [ctypes.POINTER(vp), vp, i64p, dblp, dblp, dbl, vp]),
        "fastcheb_multi_create": (i32, [ctypes.POINTER(vp), i32, ctypes.POINTER(i32), i32, i64p, i32, i32, i32, vp, i32, dblp, dblp]),
        "fastcheb_multi_destroy": (i32, [vp]),
        "fastcheb_multi_ngpus": (i32, [vp]),
        "fastcheb_multi_device": (i32, [vp, i32]),
        "fastcheb_multi_poly": (vp, [vp, i32]),
        "fastcheb_multi_eval": (i32, [vp, vp, i64, vp, i32]),
        "fastcheb_multi_eval_jac": (i32, [vp, vp, i64, vp, vp, i32]),
        "fastcheb_multi_eval_grad": (i32, [vp, vp, i64, vp, vp, i32]),
        "fastcheb_multi_eval_jac_tuple": (i32, [vp, vp, i64, vp, i32]),
        "fastcheb_multi_fit": (i32, [i32, ctypes.POINTER(i32), i32, i64p, i32, i32, i32, i64, vp, vp, vp, i32]),
        "fastcheb_peer_alloc": (i32, [ctypes.POINTER(vp), ctypes.c_size_t, vp, i32]),
        "fastcheb_peer_free": (i32, [vp, i32]),
        "fastcheb_peer_open": (i32, [ctypes.POINTER(vp), vp, i32]),
        "fastcheb_peer_close": (i32, [vp, i32]),
        "fastcheb_peer_read": (i32, [vp, vp, ctypes.c_size_t, i32]),
        "fastcheb_eval_many": (i32, [i64, i64, i32, i32, i32, vp, vp, vp, i32, vp, i64, vp, vp, vp, i32, i32, vp]),
        "fastcheb_eval_many_shared": (i32, [i64, i64, i32, i32, i32, vp, vp, vp, vp, i64, vp, vp, vp, i32, i32, vp]),
        "fastcheb_vandermonde": (i32, [i32, i64p, i32, vp, i64, dblp, dblp, vp, i64, i32, i32, vp]),
        "fastcheb_eval_vjp": (i32, [vp, vp, i64, vp, vp, vp, i32, vp]),
        "fastcheb_eval_jvp": (i32, [vp, vp, i64, vp, vp, vp, i32, vp]),
        "fastcheb_fit_async": (i32, [i32, i64p, i32, i32, i32, i64, vp, vp, vp, vp, i32, vp]),
        "fastcheb_droptol_shape": (i32, [i32, i64p, i32, i32, i32, vp, dbl, i64p, i32, i32, vp]),
        "fastcheb_interp": (i32, [ctypes.POINTER(vp), i32, i64p, i32, i32, i32, vp, i32, dblp, dblp, dbl, i32, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


def last_error() -> str:
    buf = ctypes.create_string_buffer(1024)
    lib().fastcheb_last_error(buf, len(buf))
    return buf.value.decode("utf-8", "replace")


def error_index() -> int:
    v = ctypes.c_int64(-1)
    lib().fastcheb_error_index(ctypes.byref(v))
    return v.value


def i64_array(vals):
    return (ctypes.c_int64 * len(vals))(*[int(v) for v in vals])


def dbl_array(vals):
    return (ctypes.c_double * len(vals))(*[float(v) for v in vals])

"""CPU: the C-ABI library loads without a GPU and exports every function include/fastcheb.h declares; the
host-side argument checks that need no device behave as documented (no compute call is made here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "fastcheb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fastcheb_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from fastcheb import _capi

    names = declared_functions()
    assert len(names) >= 26
    L = _capi.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(_capi.SYMBOLS) == names          # the ctypes table and the header agree
    assert L.fastcheb_version() == 100


def test_host_side_argument_checks_need_no_gpu():
    from fastcheb import _capi

    L = _capi.lib()
    h = ctypes.c_void_p()
    dims = _capi.i64_array((4,))
    c = np.zeros(4)
    lb, ub = _capi.dbl_array([0.0]), _capi.dbl_array([1.0])
    # bad rank / dtype / component count are rejected before any CUDA call
    assert L.fastcheb_create(ctypes.byref(h), 0, dims, _capi.F64, 0, 1, c.ctypes.data, _capi.MEM_HOST, lb, ub, 0) == _capi.EINVAL
    assert L.fastcheb_create(ctypes.byref(h), 1, dims, 7, 0, 1, c.ctypes.data, _capi.MEM_HOST, lb, ub, 0) == _capi.EINVAL
    assert L.fastcheb_create(ctypes.byref(h), 1, dims, _capi.F64, 0, 0, c.ctypes.data, _capi.MEM_HOST, lb, ub, 0) == _capi.EINVAL
    assert "ncomp" in _capi.last_error()
    assert L.fastcheb_destroy(None) == _capi.OK     # NULL is a no-op
    assert L.fastcheb_ndim(None) == -1
    # without a usable device every compute entry point fails loudly (no CPU fallback)
    if L.fastcheb_device_count() == 0:
        rc = L.fastcheb_create(ctypes.byref(h), 1, dims, _capi.F64, 0, 1, c.ctypes.data, _capi.MEM_HOST, lb, ub, 0)
        assert rc == _capi.ECUDA
        out = np.zeros(4)
        rc = L.fastcheb_fit(1, dims, _capi.F64, 0, 1, 1, c.ctypes.data, out.ctypes.data, None, _capi.MEM_HOST, 0, None)
        assert rc == _capi.ECUDA


def test_package_has_no_oracle_dependency():
    """The product never imports, links or loads anything under oracle/ (the oracle is the checker only)."""
    pkg = os.path.join(ROOT, "fastchebinterp.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "cheb_oracle" not in text and "oracle/" not in text.replace("vs the oracle", ""), (dirpath, f)

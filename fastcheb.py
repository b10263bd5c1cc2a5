"""Import shim: the product package lives in the directory `fastchebinterp.jl_b200/` (the name the build
contract fixes), which is not a valid Python identifier; this module loads it under the name `fastcheb`.

    import fastcheb
    c = fastcheb.chebinterp(vals, lb, ub)
"""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fastchebinterp.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "fastcheb", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["fastcheb"] = _mod
_spec.loader.exec_module(_mod)

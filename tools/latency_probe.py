import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, fastcheb
rng = np.random.default_rng(0)
def t(f, n=20):
    f(); 
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6
v1 = rng.uniform(-1, 1, 257); v2 = rng.uniform(-1, 1, (65, 65)); v3 = rng.uniform(-1, 1, (33, 33, 33, 3)); v4 = rng.uniform(-1,1,10001)
print("chebinterp 1-D n=257      %8.1f us" % t(lambda: fastcheb.chebinterp(v1, 0, 1, tol=0.0).close()))
print("chebinterp 1-D n=10001    %8.1f us" % t(lambda: fastcheb.chebinterp(v4, 0, 1, tol=0.0).close(), 5))
print("chebinterp 2-D 65x65      %8.1f us" % t(lambda: fastcheb.chebinterp(v2, [0, 0], [1, 1], tol=0.0).close()))
print("chebinterp 3-D 33^3 K=3   %8.1f us" % t(lambda: fastcheb.chebinterp(v3, [0, 0, 0], [1, 1, 1], tol=0.0).close(), 5))
print("chebcoefs  2-D 65x65      %8.1f us" % t(lambda: fastcheb.chebcoefs(v2)))
c = fastcheb.chebinterp(v2, [0, 0], [1, 1], tol=0.0)
x1 = np.array([0.3, 0.4])
print("c(x) single point         %8.1f us" % t(lambda: c(x1), 200))
xs = rng.uniform(0, 1, (1000, 2))
print("c.(xs) 1000 points        %8.1f us" % t(lambda: c(xs), 200))
cc = fastcheb.chebinterp(np.cos, 10**4, 0, 1000)
t0 = time.perf_counter(); r = fastcheb.roots(cc); print("roots(order 1e4 cos): %d roots in %.1f ms" % (len(r), (time.perf_counter() - t0) * 1e3))
t0 = time.perf_counter(); r = fastcheb.roots(cc); print("roots again: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
# the C ABI alone (ctypes call overhead ~2 us): what a Julia ccall would see
import ctypes
from fastcheb import _capi
L = _capi.lib()
flat = np.ascontiguousarray(v2.reshape(-1, order="F")); out = np.empty_like(flat)
d2 = _capi.i64_array((65, 65)); lb2, ub2 = _capi.dbl_array([0, 0]), _capi.dbl_array([1, 1])
def c_fit(): assert L.fastcheb_fit(2, d2, _capi.F64, 0, 1, 1, flat.ctypes.data, out.ctypes.data, None, _capi.MEM_HOST, 0, None) == 0
def c_interp(tol):
    h = ctypes.c_void_p()
    assert L.fastcheb_interp(ctypes.byref(h), 2, d2, _capi.F64, 0, 1, flat.ctypes.data, _capi.MEM_HOST, lb2, ub2, tol, 0, None) == 0
    return h
def c_interp_destroy(tol=0.0): L.fastcheb_destroy(c_interp(tol))
print("C ABI fastcheb_fit 65x65 (host buffers)        %8.1f us" % t(c_fit, 200))
print("C ABI fastcheb_interp + destroy 65x65, tol=0   %8.1f us" % t(lambda: c_interp_destroy(0.0), 200))
print("C ABI fastcheb_interp + destroy 65x65, tol=eps %8.1f us" % t(lambda: c_interp_destroy(-1.0), 200))
hs = []
t0 = time.perf_counter()
for _ in range(200): hs.append(c_interp(0.0))
print("C ABI fastcheb_interp alone 65x65, tol=0       %8.1f us" % ((time.perf_counter() - t0) / 200 * 1e6))
for h in hs: L.fastcheb_destroy(h)
import torch
dv = torch.from_numpy(flat).cuda(); do = torch.empty_like(dv); torch.cuda.synchronize()
sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
def c_fit_dev():
    assert L.fastcheb_fit_async(2, d2, _capi.F64, 0, 1, 1, dv.data_ptr(), do.data_ptr(), None, None, 0, sp) == 0
for _ in range(10): c_fit_dev()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): c_fit_dev()
e1.record(); torch.cuda.synchronize()
print("fit kernel alone 65x65 (device, back-to-back)  %8.1f us" % (e0.elapsed_time(e1) * 10))

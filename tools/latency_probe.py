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

import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, fastcheb
rng = np.random.default_rng(0)
def t(f, n=10):
    f()
    t0 = time.perf_counter()
    for _ in range(n): f()
    return (time.perf_counter() - t0) / n * 1e6
v3 = rng.uniform(-1, 1, (33, 33, 33, 3)); v3s = rng.uniform(-1, 1, (33, 33, 33))
print("chebcoefs 33^3 K=3      %8.1f us" % t(lambda: fastcheb.chebcoefs(v3, ndim=3)))
print("chebcoefs 33^3 scalar   %8.1f us" % t(lambda: fastcheb.chebcoefs(v3s)))
c3 = fastcheb.chebcoefs(v3, ndim=3)
print("ChebPoly create+close K=3 %8.1f us" % t(lambda: fastcheb.ChebPoly(c3, [0,0,0],[1,1,1]).close()))
def mk():
    p = fastcheb.ChebPoly(c3, [0,0,0],[1,1,1]); p(np.array([0.5,0.5,0.5])); p.close()
print("ChebPoly create+eval+close K=3 %8.1f us" % t(mk))
print("chebinterp 33^3 K=3 tol=0  %8.1f us" % t(lambda: fastcheb.chebinterp(v3, [0,0,0],[1,1,1], tol=0.0).close()))
print("chebinterp 33^3 K=3 default tol  %8.1f us" % t(lambda: fastcheb.chebinterp(v3, [0,0,0],[1,1,1]).close()))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): fastcheb.chebinterp(v3, [0,0,0],[1,1,1], tol=0.0).close()
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(14)

import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "oracle"))
import numpy as np, fastcheb, cheb_oracle as orc
rng = np.random.default_rng(0)
for shape, hm in (((3,), None), ((5,), None), ((2,), 40), ((2,), None), ((2, 2), None)):
    try:
        v = rng.uniform(-1, 1, ((hm,) if hm else ()) + shape)
        got = fastcheb.chebcoefs(v, ndim=len(shape), howmany=hm) if hm else fastcheb.chebcoefs(v, ndim=len(shape))
        want = orc.chebcoefs(v[0] if hm else v, len(shape))
        print(shape, hm, "ok", np.abs((got[0] if hm else got) - want).max())
    except Exception as e:
        print(shape, hm, "FAIL", str(e)[:100])
        break

"""The shared-memory-resident N-d / long-line DCT-I fit (fit_nd.cuh): batches of multidimensional arrays and of lines
longer than 65 through the C ABI vs the long-double oracle, 1e-13 max|c| per coefficient (450 eps for Float32).
Shapes cover every code path of the kernel: twice-folded (4 | n-1) and once-folded dimensions, odd n-1, size-1 and
size-2 dimensions, unequal dimensions, vector / complex elements, ragged 16- / 32-line blocks, the three accumulator
size classes (n <= 65, <= 129, <= 257 with the matrices read from L2), and the fall-back when an array does not fit
shared memory."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def tol_for(dt):
    return 1e-13 if dt == np.float64 else 450 * np.finfo(np.float32).eps


ND_CASES = [  # dims, K, complex, howmany
    ((65, 65), None, False, 37),
    ((33, 33, 33), None, False, 5),       # Float64: 287 KB, does not fit shared memory -> generic kernels
    ((17, 17, 17), 3, False, 9),
    ((11, 21), 2, False, 50),             # once-folded dims (n-1 = 10, 20 not both multiples of 4)
    ((64, 63), None, False, 11),          # n-1 odd / n-1 = 2 mod 4
    ((9, 1, 7), None, False, 20),
    ((2, 3, 2, 5), 2, True, 13),
    ((17, 17, 17, 17), None, False, 2),   # Float32: 334 KB -> generic; covered for parity of the fall-back
    ((33, 9), None, True, 40),
    ((5, 65, 3), None, False, 7),
    ((129, 17), None, False, 6),          # class 1 (n <= 129) inside an N-d array
    ((33, 33, 33), 2, False, 3),          # too large for shared memory in both precisions: slabs + strided tiles, vector elements
    ((17, 17, 17, 9), None, True, 3),     # ... complex, unequal last dimension
    ((65, 65, 20), None, False, 2),       # ... once-folded last dimension (n - 1 = 19)
    ((17, 17), None, False, 2501),        # small arrays in batches: a CTA takes blocks of several arrays (+ a ragged tail)
    ((33, 33), None, False, 5003),
    ((9, 9, 9), 2, False, 3000),
    ((17, 9), None, True, 2400),
]


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("case", ND_CASES, ids=lambda c: "x".join(map(str, c[0])) + f"-K{c[1]}-{'c' if c[2] else 'r'}")
def test_batched_nd_fits(fc, orc, case, dt):
    dims, K, cplx, howmany = case
    rng = np.random.default_rng(sum(dims) + howmany)
    shape = (howmany,) + dims + ((K,) if K else ())
    vals = rng.uniform(-1, 1, shape)
    if cplx:
        vals = (vals + 1j * rng.uniform(-1, 1, shape)).astype(np.complex128 if dt == np.float64 else np.complex64)
    else:
        vals = vals.astype(dt)
    got, mx = fc.chebcoefs(vals, ndim=len(dims), howmany=howmany, return_maxabs=True)
    assert got.shape == vals.shape and got.dtype == vals.dtype
    wide = vals.astype(np.complex128 if cplx else np.float64)
    for i in sorted({0, 1, howmany // 2, howmany - 1}):
        want = orc.chebcoefs(wide[i], len(dims))
        assert np.abs(got[i] - want).max() <= tol_for(dt) * np.abs(want).max(), i
    # per-array maximum of the element infnorm (interp.jl:74-78): complex modulus, max over the K components
    a = np.abs(got.astype(np.complex128 if cplx else np.float64)).reshape(howmany, -1)
    assert np.allclose(mx, a.max(axis=1), rtol=1e-6 if dt == np.float32 else 1e-13, atol=0)
    # a single array through the same entry (chebinterp's fit) equals its slot of the batch bit for bit
    one = fc.chebcoefs(vals[howmany // 2], ndim=len(dims))
    assert np.array_equal(one, got[howmany // 2])


@pytest.mark.parametrize("dt", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("n,howmany", [(129, 1000), (129, 37), (201, 333), (257, 150), (97, 200), (81, 64), (257, 1)])
def test_batched_long_lines(fc, orc, n, howmany, dt):
    rng = np.random.default_rng(n + howmany)
    vals = rng.uniform(-1, 1, (howmany, n)).astype(dt)
    got = fc.chebcoefs(vals, ndim=1, howmany=howmany)
    for i in sorted({0, howmany // 3, howmany - 1}):
        want = orc.chebcoefs(vals[i].astype(np.float64), 1)
        assert np.abs(got[i] - want).max() <= tol_for(dt) * np.abs(want).max(), (n, i)
    if howmany > 1:
        cv = (rng.uniform(-1, 1, (howmany, n, 2)) + 1j * rng.uniform(-1, 1, (howmany, n, 2))).astype(np.complex128 if dt == np.float64 else np.complex64)
        gc = fc.chebcoefs(cv, ndim=1, howmany=howmany)
        want = orc.chebcoefs(cv[howmany - 1].astype(np.complex128), 1)
        assert np.abs(gc[howmany - 1] - want).max() <= tol_for(dt) * np.abs(want).max()


def test_nd_fit_nonfinite_index_and_in_place(fc):
    vals = np.ones((300, 17, 9))
    vals[123, 5, 7] = np.nan
    vals[250, 0, 0] = np.inf
    with pytest.raises(fc.DomainError):
        fc.chebcoefs(vals, ndim=2, howmany=300)
    # Julia order inside an array: element (i1, i2) of array f is at flat index f*153 + i1 + 17*i2
    assert fc._capi.error_index() == 123 * 153 + 5 + 17 * 7
    lines = np.ones((500, 129))
    lines[321, 77] = -np.inf
    with pytest.raises(fc.DomainError):
        fc.chebcoefs(lines, ndim=1, howmany=500)
    assert fc._capi.error_index() == 321 * 129 + 77
    # vals and coefs may alias (include/fastcheb.h)
    import ctypes

    rng = np.random.default_rng(1)
    a = rng.uniform(-1, 1, (40, 33 * 33))
    ref = fc.chebcoefs(a.reshape(40, 33, 33).transpose(0, 2, 1).copy(), ndim=2, howmany=40)   # numpy (i1, i2) -> Julia order
    flat = np.ascontiguousarray(a)
    rc = fc._capi.lib().fastcheb_fit(2, fc._capi.i64_array((33, 33)), fc._capi.F64, 0, 1, 40, flat.ctypes.data, flat.ctypes.data, None,
                                     fc._capi.MEM_HOST, 0, None)
    assert rc == 0
    assert np.array_equal(flat.reshape(40, 33, 33).transpose(0, 2, 1), ref)


def test_large_array_nonfinite_index(fc):
    """An array that goes through the slab + tile path reports the first non-finite input like any other."""
    vals = np.ones((2, 33, 33, 33))
    vals[1, 5, 7, 30] = np.nan
    vals[1, 9, 9, 31] = np.inf
    with pytest.raises(fc.DomainError):
        fc.chebcoefs(vals, ndim=3, howmany=2)
    assert fc._capi.error_index() == 33 ** 3 + 5 + 33 * 7 + 33 * 33 * 30


def test_nd_fit_properties_at_size(fc):
    """Size-independent properties on a large batch: linearity, and tensor products of T_k on the grid give unit
    coefficient tensors."""
    rng = np.random.default_rng(9)
    M = 2048
    a, b = rng.uniform(-1, 1, (M, 33, 33)), rng.uniform(-1, 1, (M, 33, 33))
    ca, cb, cab = (fc.chebcoefs(v, ndim=2, howmany=M) for v in (a, b, 3.0 * a + 0.25 * b))
    assert np.abs(cab - (3.0 * ca + 0.25 * cb)).max() <= 1e-13 * np.abs(cab).max()
    j = np.arange(33)
    grid = np.arccos(np.cos(np.pi * j / 32))
    k1, k2 = rng.integers(0, 33, M), rng.integers(0, 33, M)
    t = np.cos(k1[:, None, None] * grid[None, :, None]) * np.cos(k2[:, None, None] * grid[None, None, :])
    c = fc.chebcoefs(t, ndim=2, howmany=M)
    want = np.zeros((M, 33, 33))
    want[np.arange(M), k1, k2] = 1.0
    assert np.abs(c - want).max() <= 1e-13


@pytest.mark.parametrize("env", [{"FASTCHEB_FITND_PIPE": "1"}, {"FASTCHEB_FITND_SPEC": "65"}, {"FASTCHEB_FITND_SPEC": "0"}],
                         ids=["pipeline", "spec65", "nospec"])
def test_experimental_fit_nd_variants_agree(env):
    """The opt-in variants of the resident fit kernel (warp-specialised pipeline; register-resident matrices for n = 65;
    no register-resident matrices at all) compute the same transform: run in a subprocess (the switches are read once),
    a batch large enough for the pipeline (>= 2 arrays per SM), against the oracle."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "oracle"))
import numpy as np, fastcheb, cheb_oracle as orc
rng = np.random.default_rng(5)
for dims, hm in (((65, 65), 400), ((33, 33), 700), ((17, 17, 17), 320)):
    for dt, tol in ((np.float64, 1e-13), (np.float32, 450 * np.finfo(np.float32).eps)):
        v = rng.uniform(-1, 1, (hm,) + dims).astype(dt)
        got, mx = fastcheb.chebcoefs(v, ndim=len(dims), howmany=hm, return_maxabs=True)
        for i in (0, hm // 2, hm - 1):
            want = orc.chebcoefs(v[i].astype(np.float64), len(dims))
            assert np.abs(got[i] - want).max() <= tol * np.abs(want).max(), (dims, dt, i)
        assert np.allclose(mx, np.abs(got.astype(np.float64)).reshape(hm, -1).max(axis=1), rtol=1e-6)
    bad = rng.uniform(-1, 1, (hm,) + dims)
    bad[hm - 3].flat[7] = np.nan
    try:
        fastcheb.chebcoefs(bad, ndim=len(dims), howmany=hm)
        raise SystemExit("no DomainError")
    except fastcheb.DomainError:
        pass
print("ok")
""" % (root, root)
    r = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-500:] + r.stderr[-1500:]

"""The reference's own differential property, replayed on CPU.

test/Main.hs:28-52 asserts `native == netlib BLAS 3.7.0` bit for bit over random inputs.  Here
`netlib` is the C restatement of the published Fortran (oracle/netlib_l1.c) and `native` is, in turn,
  * "restatement": the C restatement of the Haskell code (oracle/lblas_oracle.c), and
  * "reference-source": the reference's Haskell SOURCE itself, evaluated by oracle/hs_eval.py (build container
    only: /root/reference does not exist on the GPU box; fewer cases, it is an interpreter);
domains follow Main.hs row by row.
"""
import ctypes as C
import os

import numpy as np
import pytest

from tests.gen import gen_nice_float, nice_scalar, span

CASES = 300


@pytest.fixture(scope="module", params=["restatement", "reference-source"])
def oracle(request):
    """shadows conftest's fixture: every property below runs once per `native`"""
    global CASES
    from oracle import oracle as o
    o.build()
    if request.param == "restatement":
        yield o
        return
    from oracle import hs_eval
    if not os.path.isdir(hs_eval.REFERENCE_SRC):
        pytest.skip("/root/reference is only present in the build container")
    from tests.ref_adapter import ReferenceAsOracle
    saved, CASES = CASES, 120
    try:
        yield ReferenceAsOracle()
    finally:
        CASES = saved
DTYPES = [np.float32, np.float64]            # the reference runs every property for Float and Double (Main.hs:28-52)


def p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double if a.dtype == np.float64 else C.c_float))


def bits(a, dtype=None):
    a = np.asarray(a) if dtype is None else np.asarray(a, dtype=dtype)
    if a.dtype not in (np.float32, np.float64):
        a = a.astype(np.float32)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def same(a, b, dtype=None):
    if dtype is None:
        dtype = np.asarray(a).dtype if np.asarray(a).dtype in (np.float32, np.float64) else np.float32
    return np.array_equal(bits(a, dtype), bits(b, dtype))


def eq(a, b, dtype=np.float32):
    """Haskell's Eq on Float/Double, which is what runTest uses (Main.hs:262-270): -0 == 0, NaN /= NaN."""
    return np.array_equal(np.asarray(a, dtype=dtype), np.asarray(b, dtype=dtype))


def nl(oracle, dtype, name):
    pre = "d" if dtype == np.float64 else "s"
    if name == "iamax":
        return getattr(oracle.netlib(), "nl_i" + pre + "amax")
    return getattr(oracle.netlib(), "nl_" + pre + name)


def test_readme_worked_example(oracle):
    # README.md:57-59: sdot 3 [1,2,3] 1 [4,5,6] 1 == 32
    assert oracle.sdot(3, np.array([1, 2, 3], np.float32), 1, np.array([4, 5, 6], np.float32), 1) == np.float32(32)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("seed", range(3))
def test_dot_property(oracle, seed, dtype):   # Main.hs:28-29,59-79
    rng = np.random.default_rng(100 + seed)
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); incx, incy = (int(v) for v in rng.integers(-5, 6, 2))
        u = gen_nice_float(rng, span(n, incx), dtype); v = gen_nice_float(rng, span(n, incy), dtype)
        assert same(oracle.dot(n, u, incx, v, incy), dtype(nl(oracle, dtype, "dot")(n, p(u), incx, p(v), incy)))


def test_sdsdot_property(oracle):             # Main.hs:30,84-102
    rng, nl = np.random.default_rng(7), oracle.netlib()  # noqa: F811
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); a = nice_scalar(rng)
        incx, incy = (int(v) for v in rng.integers(-5, 6, 2))
        u = gen_nice_float(rng, span(n, incx)); v = gen_nice_float(rng, span(n, incy))
        assert same(oracle.sdsdot(n, a, u, incx, v, incy), nl.nl_sdsdot(n, float(a), p(u), incx, p(v), incy))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["asum", "nrm2", "iamax"])
def test_ivi_property(oracle, name, dtype):   # Main.hs:37-42,239-257 (n from 0, incx in 1..5)
    rng = np.random.default_rng(11)
    for _ in range(CASES):
        n = int(rng.integers(0, 101)); incx = int(rng.integers(1, 6))
        u = gen_nice_float(rng, max(1 + (n - 1) * incx, 0), dtype)
        want = nl(oracle, dtype, name)(n, p(u) if u.size else None, incx)
        if name == "iamax":                   # compared as succ (iamax ..), Main.hs:41
            assert oracle.iamax(n, u, incx) + 1 == want
        else:
            assert same(getattr(oracle, name)(n, u, incx), dtype(want))


@pytest.mark.parametrize("dtype", DTYPES)
def test_scal_property(oracle, dtype):        # Main.hs:43-44,405-424
    rng = np.random.default_rng(13)
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); a = nice_scalar(rng, dtype); incx = int(rng.integers(1, 6))
        u = gen_nice_float(rng, span(n, incx), dtype); w = u.copy()
        nl(oracle, dtype, "scal")(n, float(a), p(w), incx)
        assert same(oracle.scal(n, a, u, incx), w)


@pytest.mark.parametrize("dtype", DTYPES)
def test_copy_property(oracle, dtype):        # Main.hs:45-46,429-451 (inc 0 included, tails kept)
    rng = np.random.default_rng(17)
    for _ in range(CASES * 3):
        n = int(rng.integers(1, 6)); mx, my = (int(v) for v in rng.integers(1, 3, 2))
        incx, incy = (int(v) for v in rng.integers(-5, 6, 2))
        u = gen_nice_float(rng, span(n, incx, mx), dtype); v = gen_nice_float(rng, span(n, incy, my), dtype); w = v.copy()
        nl(oracle, dtype, "copy")(n, p(u), incx, p(w), incy)
        assert same(oracle.copy(n, u, incx, v, incy), w)


def _nz(rng, lo, hi):
    while True:
        v = int(rng.integers(lo, hi + 1))
        if v != 0:
            return v


@pytest.mark.parametrize("dtype", DTYPES)
def test_swap_property(oracle, dtype):        # Main.hs:47-48,107-129
    rng = np.random.default_rng(19)
    for _ in range(CASES * 3):
        n = int(rng.integers(1, 6)); mx, my = (int(v) for v in rng.integers(0, 3, 2))
        incx, incy = _nz(rng, -5, 5), _nz(rng, -5, 5)
        u = gen_nice_float(rng, span(n, incx, mx), dtype); v = gen_nice_float(rng, span(n, incy, my), dtype)
        a, b = u.copy(), v.copy()
        nl(oracle, dtype, "swap")(n, p(a), incx, p(b), incy)
        xo, yo = oracle.swap(n, u, incx, v, incy)
        assert same(xo, a) and same(yo, b)


@pytest.mark.parametrize("dtype", DTYPES)
def test_rot_property(oracle, dtype):         # Main.hs:49-50,153-176 (inc in {-5,5}, c=cos a, s=sin a)
    rng = np.random.default_rng(23)
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); ang = nice_scalar(rng, dtype)
        incx, incy = (int(v) for v in rng.choice([-5, 5], 2))
        c, s = dtype(np.cos(ang)), dtype(np.sin(ang))
        u = gen_nice_float(rng, span(n, incx), dtype); v = gen_nice_float(rng, span(n, incy), dtype)
        a, b = u.copy(), v.copy()
        nl(oracle, dtype, "rot")(n, p(a), incx, p(b), incy, float(c), float(s))
        xo, yo = oracle.rot(n, u, incx, v, incy, c, s)
        assert same(xo, a) and same(yo, b)


@pytest.mark.parametrize("dtype", DTYPES)
def test_axpy_property(oracle, dtype):        # Main.hs:51-52,456-477
    rng = np.random.default_rng(29)
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); a = nice_scalar(rng, dtype)
        incx, incy = (int(v) for v in rng.choice([-5, 5], 2))
        mx, my = (int(v) for v in rng.integers(1, 101, 2))
        u = gen_nice_float(rng, mx + (n - 1) * abs(incx), dtype); v = gen_nice_float(rng, my + (n - 1) * abs(incy), dtype)
        w = v.copy()
        nl(oracle, dtype, "axpy")(n, float(a), p(u), incx, p(w), incy)
        assert same(oracle.axpy(n, a, u, incx, v, incy), w)


@pytest.mark.parametrize("dtype", DTYPES)
def test_rotg_property(oracle, dtype):        # Main.hs:31-32,134-148
    rng = np.random.default_rng(31)
    CT = C.c_double if dtype == np.float64 else C.c_float
    for _ in range(CASES * 3):
        sa, sb = nice_scalar(rng, dtype), nice_scalar(rng, dtype)
        a, b, c, s = (CT(float(v)) for v in (sa, sb, 0, 0))
        nl(oracle, dtype, "rotg")(C.byref(a), C.byref(b), C.byref(c), C.byref(s))
        got = (oracle.drotg if dtype == np.float64 else oracle.srotg)(sa, sb)
        # value equality: for sa == 0 > sb netlib yields c = -0 where the reference writes the
        # literal 0 (Single.hs:289); the reference's `==` accepts that and so do we
        assert eq(list(got), [a.value, b.value, c.value, s.value], dtype)


def _netlib_rotmg(fn, dtype, sd1, sd2, sx1, sy1):
    CT = C.c_double if dtype == np.float64 else C.c_float
    d1, d2, x1 = (CT(float(v)) for v in (sd1, sd2, sx1))
    par = np.zeros(5, dtype=dtype)
    fn(C.byref(d1), C.byref(d2), C.byref(x1), float(sy1), p(par))
    flag = int(par[0])
    f32 = dtype
    # the view test/Foreign.hs:245-253 takes of the Fortran outputs
    if flag == -2:
        return (-2,)
    if flag == -1:
        return (-1, f32(d1.value), f32(d2.value), f32(x1.value), par[1], par[3], par[2], par[4])
    if flag == 0:
        return (0, f32(d1.value), f32(d2.value), f32(x1.value), par[3], par[2])
    return (1, f32(d1.value), f32(d2.value), f32(x1.value), par[1], par[4])


def _view(m):
    if m.flag == -2:
        return (-2,)
    if m.flag == -1:
        return (-1, m.d1, m.d2, m.x1, m.h11, m.h12, m.h21, m.h22)
    if m.flag == 0:
        return (0, m.d1, m.d2, m.x1, m.h12, m.h21)
    return (1, m.d1, m.d2, m.x1, m.h11, m.h22)


@pytest.mark.parametrize("dtype", DTYPES)
def test_rotmg_property(oracle, dtype):       # Main.hs:33-34,181-197
    rng = np.random.default_rng(37)
    seen, skipped = set(), 0
    rotmg = oracle.drotmg if dtype == np.float64 else oracle.srotmg
    for it in range(CASES * 12):
        # the reference's generator; plus moderate magnitudes, which at Double are the only way to FLAG0 / FLAG1
        args = [nice_scalar(rng, dtype) for _ in range(4)] if it < CASES * 10 else list(rng.uniform(-4, 4, 4).astype(dtype))
        want, got = _netlib_rotmg(nl(oracle, dtype, "rotmg"), dtype, *args), _view(rotmg(*args))
        if dtype == np.float64 and want != got:
            # drotmg.f carries GAMSQ = 16777216, RGAMSQ = 5.9604645e-8 where the polymorphic reference uses the
            # single-precision literals 1.67772e7, 5.96046e-8 (Single.hs:367-369) for both instances: a d1 or d2
            # that lands between the two thresholds is rescaled by one and not by the other.  The reference's own
            # drotmg test has the same (rare) exposure; such cases are counted, not compared.
            mags = [abs(float(v)) for v in list(want[1:3]) + list(got[1:3]) if float(v) != 0]
            if any(1.67772e7 <= m * s <= 16777216.0 * 1.0000001 or 5.96046e-8 * 0.9999999 <= m * s <= 5.9604645e-8
                   for m in mags for s in (1.0, 4096.0 ** 2, 4096.0 ** -2)):
                skipped += 1
                continue
        seen.add(got[0])
        assert want[0] == got[0] and same(list(want[1:]), list(got[1:]), dtype), (args, want, got)
    assert seen == {-2, -1, 0, 1}
    assert skipped <= 3


@pytest.mark.parametrize("dtype", DTYPES)
def test_rotm_property(oracle, dtype):        # Main.hs:35-36,202-225 (flags built by rotmg)
    rng = np.random.default_rng(41)
    seen = set()
    rotmg = oracle.drotmg if dtype == np.float64 else oracle.srotmg
    for it in range(CASES * 2):
        args = [nice_scalar(rng, dtype) for _ in range(4)] if it % 4 else list(rng.uniform(-4, 4, 4).astype(dtype))
        flags = rotmg(*args)
        seen.add(flags.flag)
        n = int(rng.integers(1, 101)); incx, incy = _nz(rng, -5, 5), _nz(rng, -5, 5)
        mx, my = (int(v) for v in rng.integers(1, 101, 2))
        u = gen_nice_float(rng, mx + (n - 1) * abs(incx), dtype); v = gen_nice_float(rng, my + (n - 1) * abs(incy), dtype)
        a, b = u.copy(), v.copy()
        par = flags.sparam(dtype)
        nl(oracle, dtype, "rotm")(n, p(a), incx, p(b), incy, p(par))
        xo, yo = oracle.rotm(flags, n, u, incx, v, incy)
        assert same(xo, a) and same(yo, b)
    assert seen == {-2, -1, 0, 1}

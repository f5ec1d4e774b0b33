"""Parity of the CUDA path (through the C ABI) with the oracle, property by property.

Each test re-states one property of the reference's suite (test/Main.hs:28-52) with the same
n / increment / vector-length domains and the genNiceFloat distribution (test/Gen.hs:39-44), with
the oracle (oracle/lblas_oracle.c, the restatement of the Haskell code) where the reference puts
Fortran BLAS.  Bars: isamax bit-exact; elementwise routines bit-exact (stricter than the <= 1 ulp
the north star asks for); reductions within K*n*eps*sum|x_i*y_i| with K = 2, eps = 2^-24.
"""
import numpy as np
import pytest

from tests.gen import gen_nice_float, nice_scalar, span

pytestmark = pytest.mark.gpu

CASES = 120
K = 2.0
DTYPES = [np.float32, np.float64]      # the reference runs every property at Float and at Double (Main.hs:28-52)


def eps_of(dtype):
    return np.float64(2.0 ** -24 if np.dtype(dtype) == np.float32 else 2.0 ** -53)


def bits(a):
    a = np.asarray(a)
    if a.dtype != np.float64:
        a = a.astype(np.float32)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def assert_same(got, want, what=""):
    """bit-identical, except that any NaN matches any NaN (payloads differ between x86 and the GPU)."""
    want = np.asarray(want)
    dt = np.float64 if want.dtype == np.float64 else np.float32
    got, want = np.asarray(got, dtype=dt), np.asarray(want, dtype=dt)
    assert got.shape == want.shape, what
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan), what
    assert np.array_equal(bits(got)[~nan], bits(want)[~nan]), f"{what}: {got} vs {want}"


def nz(rng, lo, hi):
    while True:
        v = int(rng.integers(lo, hi + 1))
        if v != 0:
            return v


def sample(v, n, inc):
    first = 0 if inc > 0 else (1 - n) * inc
    return np.asarray(v, dtype=np.float64)[first + np.arange(n) * inc]


def test_readme_example(blas, ctx):
    # README.md:57-59
    u, v = ctx.to_device([1, 2, 3]), ctx.to_device([4, 5, 6])
    assert blas.sdot(3, u, 1, v, 1) == np.float32(32)


@pytest.mark.parametrize("dtype", DTYPES)
def test_dot(blas, ctx, oracle, dtype):              # Main.hs:28,59-79: n 1..100, inc in [-5..5] incl. 0
    rng = np.random.default_rng(1)
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); incx, incy = (int(v) for v in rng.integers(-5, 6, 2))
        u = gen_nice_float(rng, span(n, incx), dtype); v = gen_nice_float(rng, span(n, incy), dtype)
        got = blas.sdot(n, ctx.to_device(u), incx, ctx.to_device(v), incy)
        want = oracle.sdot(n, u, incx, v, incy)
        bound = K * n * eps_of(dtype) * np.sum(np.abs(sample(u, n, incx) * sample(v, n, incy)))
        assert abs(np.float64(got) - np.float64(want)) <= bound, (n, incx, incy, got, want)


def test_sdsdot(blas, ctx, oracle):           # Main.hs:30,84-102
    dtype = np.float32
    rng = np.random.default_rng(2)
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); a = nice_scalar(rng, dtype)
        incx, incy = (int(v) for v in rng.integers(-5, 6, 2))
        u = gen_nice_float(rng, span(n, incx), dtype); v = gen_nice_float(rng, span(n, incy), dtype)
        got = blas.sdsdot(n, a, ctx.to_device(u), incx, ctx.to_device(v), incy)
        want = oracle.sdsdot(n, a, u, incx, v, incy)
        # double accumulation: the only differences are the order of exact-product sums in double
        # and one final rounding -> within 1 ulp of the float result
        assert abs(np.float64(got) - np.float64(want)) <= np.spacing(np.abs(want)), (n, incx, incy, got, want)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("name", ["sasum", "snrm2", "isamax"])
def test_ivi(blas, ctx, oracle, name, dtype):        # Main.hs:37-42,239-257: n 0..100, incx 1..5
    rng = np.random.default_rng(3)
    for _ in range(CASES):
        n = int(rng.integers(0, 101)); incx = int(rng.integers(1, 6))
        u = gen_nice_float(rng, max(1 + (n - 1) * incx, 0), dtype)
        got = getattr(blas, name)(n, ctx.to_device(u), incx)
        want = getattr(oracle, name)(n, u, incx)
        if name == "isamax":
            assert got == want, (n, incx, u)
        elif name == "sasum":
            assert abs(np.float64(got) - np.float64(want)) <= K * max(n, 1) * eps_of(dtype) * np.sum(np.abs(sample(u, n, incx))) if n else got == want
        else:
            assert abs(np.float64(got) - np.float64(want)) <= K * max(n, 1) * eps_of(dtype) * np.float64(want), (n, incx, got, want)


@pytest.mark.parametrize("dtype", DTYPES)
def test_scal(blas, ctx, oracle, dtype):             # Main.hs:43,405-424
    rng = np.random.default_rng(4)
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); a = nice_scalar(rng, dtype); incx = int(rng.integers(1, 6))
        u = gen_nice_float(rng, span(n, incx), dtype)
        du = ctx.to_device(u)
        assert_same(blas.sscal(n, a, du, incx).to_host(), oracle.sscal(n, a, u, incx))
        assert_same(du.to_host(), u, "input untouched")


@pytest.mark.parametrize("dtype", DTYPES)
def test_copy(blas, ctx, oracle, dtype):             # Main.hs:45,429-451: n 1..5, inc incl. 0, tails survive
    rng = np.random.default_rng(5)
    for _ in range(CASES * 2):
        n = int(rng.integers(1, 6)); mx, my = (int(v) for v in rng.integers(1, 3, 2))
        incx, incy = (int(v) for v in rng.integers(-5, 6, 2))
        u = gen_nice_float(rng, span(n, incx, mx), dtype); v = gen_nice_float(rng, span(n, incy, my), dtype)
        got = blas.scopy(n, ctx.to_device(u), incx, ctx.to_device(v), incy).to_host()
        assert_same(got, oracle.scopy(n, u, incx, v, incy), (n, incx, incy))


@pytest.mark.parametrize("dtype", DTYPES)
def test_swap(blas, ctx, oracle, dtype):             # Main.hs:47,107-129
    rng = np.random.default_rng(6)
    for _ in range(CASES * 2):
        n = int(rng.integers(1, 6)); mx, my = (int(v) for v in rng.integers(0, 3, 2))
        incx, incy = nz(rng, -5, 5), nz(rng, -5, 5)
        u = gen_nice_float(rng, span(n, incx, mx), dtype); v = gen_nice_float(rng, span(n, incy, my), dtype)
        xo, yo = blas.sswap(n, ctx.to_device(u), incx, ctx.to_device(v), incy)
        wx, wy = oracle.sswap(n, u, incx, v, incy)
        assert_same(xo.to_host(), wx); assert_same(yo.to_host(), wy)


@pytest.mark.parametrize("dtype", DTYPES)
def test_rot(blas, ctx, oracle, dtype):              # Main.hs:49,153-176: inc in {-5,5}
    rng = np.random.default_rng(7)
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); ang = nice_scalar(rng, dtype)
        incx, incy = (int(v) for v in rng.choice([-5, 5], 2))
        c, s = dtype(np.cos(ang)), dtype(np.sin(ang))
        u = gen_nice_float(rng, span(n, incx), dtype); v = gen_nice_float(rng, span(n, incy), dtype)
        xo, yo = blas.srot(n, ctx.to_device(u), incx, ctx.to_device(v), incy, c, s)
        wx, wy = oracle.srot(n, u, incx, v, incy, c, s)
        assert_same(xo.to_host(), wx); assert_same(yo.to_host(), wy)


@pytest.mark.parametrize("dtype", DTYPES)
def test_axpy(blas, ctx, oracle, dtype):             # Main.hs:51,456-477
    rng = np.random.default_rng(8)
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); a = nice_scalar(rng, dtype)
        incx, incy = (int(v) for v in rng.choice([-5, 5], 2))
        mx, my = (int(v) for v in rng.integers(1, 101, 2))
        u = gen_nice_float(rng, mx + (n - 1) * abs(incx), dtype); v = gen_nice_float(rng, my + (n - 1) * abs(incy), dtype)
        dv = ctx.to_device(v)
        got = blas.saxpy(n, a, ctx.to_device(u), incx, dv, incy).to_host()
        assert_same(got, oracle.saxpy(n, a, u, incx, v, incy))
        assert_same(dv.to_host(), v, "input untouched")


@pytest.mark.parametrize("dtype", DTYPES)
def test_rotg_rotmg(blas, oracle, dtype):            # Main.hs:31,33 -- host scalars, bit-exact
    rng = np.random.default_rng(9)
    seen = set()
    p = "s" if dtype == np.float32 else "d"
    rotg, rotmg, orotg, orotmg = (getattr(m, p + n) for m in (blas, oracle) for n in ("rotg", "rotmg"))
    for it in range(2400):
        a = [nice_scalar(rng, dtype) for _ in range(4)] if it < 2000 else list(rng.uniform(-4, 4, 4).astype(dtype))
        assert_same(np.array(list(rotg(a[0], a[1])), dtype=dtype), np.array(list(orotg(a[0], a[1])), dtype=dtype))
        g, w = rotmg(*a), orotmg(*a)
        seen.add(g.flag)
        assert g.flag == w.flag
        assert_same(np.array([g.d1, g.d2, g.x1, g.h11, g.h12, g.h21, g.h22], dtype=dtype), np.array(list(w)[1:], dtype=dtype))
    assert seen == {-2, -1, 0, 1}


@pytest.mark.parametrize("dtype", DTYPES)
def test_rotm(blas, ctx, oracle, dtype):             # Main.hs:35,202-225: flags from rotmg, all four constructors
    rng = np.random.default_rng(10)
    seen = set()
    for _ in range(CASES * 2):
        args = [nice_scalar(rng, dtype) for _ in range(4)] if _ % 3 else list(rng.uniform(-4, 4, 4).astype(dtype))
        p = "s" if dtype == np.float32 else "d"
        flags, oflags = getattr(blas, p + "rotmg")(*args), getattr(oracle, p + "rotmg")(*args)
        seen.add(flags.flag)
        n = int(rng.integers(1, 101)); incx, incy = nz(rng, -5, 5), nz(rng, -5, 5)
        mx, my = (int(v) for v in rng.integers(1, 101, 2))
        u = gen_nice_float(rng, mx + (n - 1) * abs(incx), dtype); v = gen_nice_float(rng, my + (n - 1) * abs(incy), dtype)
        du, dv = ctx.to_device(u), ctx.to_device(v)
        xo, yo = blas.srotm(flags, n, du, incx, dv, incy)
        wx, wy = oracle.srotm(oflags, n, u, incx, v, incy)
        assert_same(xo.to_host(), wx); assert_same(yo.to_host(), wy)
        if flags.flag == -2:
            assert xo is du and yo is dv           # Single.hs:473 returns the arguments themselves
    assert seen == {-2, -1, 0, 1}

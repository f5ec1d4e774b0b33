"""The oracle against a REAL compiled BLAS: OpenBLAS through scipy.linalg.blas.

The reference's test-suite pins every routine to a compiled Fortran BLAS through FFI
(test/Main.hs:28-52, test/Foreign.hs:32-81).  That library (netlib 3.7.0 + gfortran) cannot be built
here, but scipy ships OpenBLAS behind the same Fortran interface, so the reference's differential
property can run against an independent, real implementation -- on this box and on the GPU box.
OpenBLAS is not netlib: its kernels vectorise sums and contract multiply-adds, so
  * moves, single multiplies and index results must agree BIT FOR BIT (copy swap scal iamax);
  * multiply-add routines agree within one rounding of each term (axpy rot rotm);
  * sums agree within the reduction bound n*eps*sum|x_i y_i| (dot asum nrm2 sdsdot).
Domains follow test/Main.hs row by row, minus inc == 0 (scipy's wrappers reject it).
"""
import numpy as np
import pytest

from tests.gen import gen_nice_float, nice_scalar, span

sblas = pytest.importorskip("scipy.linalg.blas")

CASES = 200
DTYPES = [np.float32, np.float64]


def fn(dtype, name):
    pre = "d" if dtype == np.float64 else "s"
    if name == "iamax":
        return getattr(sblas, "i" + pre + "amax")
    return getattr(sblas, pre + name)


def bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def nz_inc(rng, lo=-5, hi=5):
    while True:
        v = int(rng.integers(lo, hi + 1))
        if v != 0:
            return v


def eps(dtype):
    return float(np.finfo(dtype).eps) / 2      # unit round-off


def openblas_sdsdot():
    """scipy.linalg.blas has no sdsdot wrapper; the bundled OpenBLAS exports it as scipy_cblas_sdsdot."""
    import ctypes as C
    import glob
    import os
    import scipy
    for path in glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so")):
        L = C.CDLL(path)
        if hasattr(L, "scipy_cblas_sdsdot"):
            f = L.scipy_cblas_sdsdot
            f.restype = C.c_float
            f.argtypes = [C.c_int, C.c_float, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int]
            return lambda n, sb, x, incx, y, incy: f(n, float(sb), x.ctypes.data_as(C.POINTER(C.c_float)), incx,
                                                     y.ctypes.data_as(C.POINTER(C.c_float)), incy)
    return None


@pytest.mark.parametrize("dtype", DTYPES)
def test_copy_swap_scal_bit_exact(oracle, dtype):          # Main.hs:43-48
    rng = np.random.default_rng(7)
    for _ in range(CASES):
        n = int(rng.integers(1, 6)); incx, incy = nz_inc(rng), nz_inc(rng)
        x = gen_nice_float(rng, span(n, incx, int(rng.integers(0, 3))), dtype)
        y = gen_nice_float(rng, span(n, incy, int(rng.integers(0, 3))), dtype)
        want = fn(dtype, "copy")(x, y.copy(), n=n, incx=incx, incy=incy)
        assert np.array_equal(bits(oracle.copy(n, x, incx, y, incy)), bits(want))
        wx, wy = fn(dtype, "swap")(x.copy(), y.copy(), n=n, incx=incx, incy=incy)
        gx, gy = oracle.swap(n, x, incx, y, incy)
        assert np.array_equal(bits(gx), bits(wx)) and np.array_equal(bits(gy), bits(wy))
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); incx = int(rng.integers(1, 6))
        x = gen_nice_float(rng, span(n, incx), dtype); a = nice_scalar(rng, dtype)
        want = fn(dtype, "scal")(a, x.copy(), n=n, incx=incx)
        assert np.array_equal(bits(oracle.scal(n, a, x, incx)), bits(want))


@pytest.mark.parametrize("dtype", DTYPES)
def test_iamax_exact(oracle, dtype):                       # Main.hs:41-42 (the reference adds 1; scipy subtracts it back)
    rng = np.random.default_rng(8)
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); incx = int(rng.integers(1, 6))
        x = gen_nice_float(rng, span(n, incx), dtype)
        if rng.random() < 0.3:                             # force ties: the FIRST maximum must win
            x[:: max(1, n // 3)] = x[np.argmax(np.abs(x))]
        assert oracle.iamax(n, x, incx) == int(fn(dtype, "iamax")(x, n=n, incx=incx))


@pytest.mark.parametrize("dtype", DTYPES)
def test_axpy_rot_rotm_within_one_rounding(oracle, dtype):   # Main.hs:35-36,49-52
    rng = np.random.default_rng(9)
    u = eps(dtype)
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); incx, incy = nz_inc(rng), nz_inc(rng)
        x = gen_nice_float(rng, span(n, incx), dtype); y = gen_nice_float(rng, span(n, incy), dtype)
        a = nice_scalar(rng, dtype)
        got = oracle.axpy(n, a, x, incx, y, incy).astype(np.float64)
        want = fn(dtype, "axpy")(x, y.copy(), n=n, a=a, incx=incx, incy=incy).astype(np.float64)
        kx = np.arange(n) * incx + (0 if incx > 0 else (1 - n) * incx)
        ky = np.arange(n) * incy + (0 if incy > 0 else (1 - n) * incy)
        bound = np.zeros(y.size)
        np.add.at(bound, ky, 0)
        bound[ky] = 2 * u * (np.abs(y[ky].astype(np.float64)) + np.abs(float(a) * x[kx].astype(np.float64)))
        assert np.all(np.abs(got - want) <= bound + 0.0), (n, incx, incy)

        ang = float(nice_scalar(rng, dtype)); c, s = dtype(np.cos(ang)), dtype(np.sin(ang))
        gx, gy = oracle.rot(n, x, incx, y, incy, c, s)
        wx, wy = fn(dtype, "rot")(x.copy(), y.copy(), c, s, n=n, incx=incx, incy=incy)
        mag = np.zeros(max(x.size, y.size))
        mx = np.abs(x[kx].astype(np.float64)) + np.abs(y[ky].astype(np.float64))
        bx, by = np.zeros(x.size), np.zeros(y.size)
        bx[kx] = 3 * u * mx; by[ky] = 3 * u * mx
        assert np.all(np.abs(gx.astype(np.float64) - wx) <= bx) and np.all(np.abs(gy.astype(np.float64) - wy) <= by)
        del mag

        flags = oracle.srotmg(*gen_nice_float(rng, 4, dtype)) if dtype == np.float32 else oracle.drotmg(*gen_nice_float(rng, 4, dtype))
        param = flags.sparam(dtype)
        gx, gy = oracle.rotm(flags, n, x, incx, y, incy)
        wx, wy = fn(dtype, "rotm")(x.copy(), y.copy(), param, n=n, incx=incx, incy=incy)
        h = float(np.max(np.abs(param[1:]))) + 1.0
        bx[kx] = 3 * u * h * mx; by[ky] = 3 * u * h * mx
        assert np.all(np.abs(gx.astype(np.float64) - wx) <= bx) and np.all(np.abs(gy.astype(np.float64) - wy) <= by)


@pytest.mark.parametrize("dtype", DTYPES)
def test_reductions_within_bound(oracle, dtype):           # Main.hs:28-30,37-40
    rng = np.random.default_rng(10)
    u = eps(dtype)
    sdsdot = openblas_sdsdot() if dtype == np.float32 else None
    for _ in range(CASES):
        n = int(rng.integers(1, 101)); incx, incy = nz_inc(rng), nz_inc(rng)
        x = gen_nice_float(rng, span(n, incx), dtype); y = gen_nice_float(rng, span(n, incy), dtype)
        kx = np.arange(n) * incx + (0 if incx > 0 else (1 - n) * incx)
        ky = np.arange(n) * incy + (0 if incy > 0 else (1 - n) * incy)
        mag = float(np.sum(np.abs(x[kx].astype(np.float64) * y[ky].astype(np.float64))))
        assert abs(float(oracle.dot(n, x, incx, y, incy)) - float(fn(dtype, "dot")(x, y, n=n, incx=incx, incy=incy))) <= 2 * n * u * mag
        if sdsdot is not None:                              # double accumulation: one float rounding at the end
            sb = nice_scalar(rng, dtype)
            got, want = float(oracle.sdsdot(n, sb, x, incx, y, incy)), float(sdsdot(n, sb, x, incx, y, incy))
            assert abs(got - want) <= 2 * u * max(abs(got), abs(want)) + 2 * n * eps(np.float64) * mag
        ix = abs(incx)
        xs = x[:: ix][:n].astype(np.float64)
        assert abs(float(oracle.asum(n, x, ix)) - float(fn(dtype, "asum")(x, n=n, incx=ix))) <= 2 * n * u * float(np.sum(np.abs(xs)))
        want = float(fn(dtype, "nrm2")(x, n=n, incx=ix))
        assert abs(float(oracle.nrm2(n, x, ix)) - want) <= 2 * n * u * want


@pytest.mark.parametrize("dtype", DTYPES)
def test_rotg_rotmg(oracle, dtype):                        # Main.hs:31-34
    rng = np.random.default_rng(11)
    u = eps(dtype)
    rotg = oracle.srotg if dtype == np.float32 else oracle.drotg
    rotmg = oracle.srotmg if dtype == np.float32 else oracle.drotmg
    for _ in range(CASES):
        a, b = gen_nice_float(rng, 2, dtype)
        c, s = fn(dtype, "rotg")(a, b)
        g = rotg(a, b)
        assert abs(float(g.c) - float(c)) <= 8 * u and abs(float(g.s) - float(s)) <= 8 * u
    # OpenBLAS re-implements srotmg rather than porting netlib's (which Single.hs:318-401 follows and
    # tests/test_oracle_netlib.py pins bit for bit): it diverges when d1/d2 leave (rgamsq, gamsq) and for
    # zero d1.  So this is a sanity check, not a pin: where neither side rescaled (every |h| below
    # gam = 4096), d1 is non-zero and the flags agree, the matrices agree to rounding.
    compared = 0
    for _ in range(4 * CASES):
        d1, d2, x1, y1 = gen_nice_float(rng, 4, dtype)
        want = fn(dtype, "rotmg")(d1, d2, x1, y1)
        got = rotmg(d1, d2, x1, y1).sparam(dtype)
        if d1 == 0 or int(want[0]) != int(got[0]) or int(got[0]) == -2:
            continue
        live = {-1: [1, 2, 3, 4], 0: [2, 3], 1: [1, 4]}[int(got[0])]
        if any(abs(float(v[k])) >= 4096.0 for v in (got, want) for k in live):
            continue
        compared += 1
        for k in live:
            assert abs(float(got[k]) - float(want[k])) <= 16 * u * abs(float(want[k])), (d1, d2, x1, y1, got, want)
    assert compared >= CASES

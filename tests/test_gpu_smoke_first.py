"""First-contact GPU checks: unit-stride fast path at sizes that exercise head/tail peeling, every
launch variant, the in-place forms and the device data generator."""
import numpy as np
import pytest

import tests.test_gpu_parity as P
from tests.gen import gen_unit

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", P.DTYPES)
def test_fill_hash_matches_numpy(blas, ctx, dtype):
    v = ctx.alloc(100003, dtype).fill_hash(7, -1.0, 1.0)
    assert v.dtype == dtype
    P.assert_same(v.to_host(), blas.hash_fill_reference(100003, 7, -1.0, 1.0, dtype=dtype))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 8, 31, 33, 1000, 4099, 65537, 1 << 20, (1 << 20) + 3])
@pytest.mark.parametrize("off", [0, 1, 3, 5])
@pytest.mark.parametrize("dtype", P.DTYPES)
def test_unit_paths_all_alignments(blas, ctx, oracle, n, off, dtype):
    rng = np.random.default_rng(n * 8 + off)
    E = float(P.eps_of(dtype))
    u, v = gen_unit(rng, n + off, dtype), gen_unit(rng, n + off + 2, dtype)
    du, dv = ctx.to_device(u).view(off, n), ctx.to_device(v).view(off + 2 if off % 2 else off, n)
    hu, hv = du.to_host(), dv.to_host()
    got = blas.sdot(n, du, 1, dv, 1)
    want = oracle.sdot(n, hu, 1, hv, 1)
    assert abs(float(got) - float(want)) <= 2 * n * E * float(np.sum(np.abs(hu.astype(np.float64) * hv)))
    assert blas.isamax(n, du, 1) == oracle.isamax(n, hu, 1)
    assert abs(float(blas.snrm2(n, du, 1)) - float(oracle.snrm2(n, hu, 1))) <= 2 * n * E * float(oracle.snrm2(n, hu, 1))
    assert abs(float(blas.sasum(n, du, 1)) - float(oracle.sasum(n, hu, 1))) <= 2 * n * E * float(np.sum(np.abs(hu)))
    P.assert_same(blas.saxpy(n, 1.000123, du, 1, dv, 1).to_host(), oracle.saxpy(n, 1.000123, hu, 1, hv, 1))
    xo, yo = blas.srot(n, du, 1, dv, 1, np.cos(0.3), np.sin(0.3))
    wx, wy = oracle.srot(n, hu, 1, hv, 1, np.cos(0.3), np.sin(0.3))
    P.assert_same(xo.to_host(), wx); P.assert_same(yo.to_host(), wy)


@pytest.mark.parametrize("variant", range(23))     # every (pack, unroll, policy) shape, the load-fenced ones included
@pytest.mark.parametrize("threads", [128, 512])
@pytest.mark.parametrize("dtype", P.DTYPES)
def test_every_variant(blas, ctx, oracle, variant, threads, dtype):
    n = 300007
    rng = np.random.default_rng(variant)
    E = float(P.eps_of(dtype))
    u, v = gen_unit(rng, n, dtype), gen_unit(rng, n, dtype)
    u[12345] = -3.0; u[99999] = 3.0          # tie: the first wins
    du, dv = ctx.to_device(u), ctx.to_device(v)
    ctx.set_launch("all", variant, threads, 0)
    try:
        assert blas.isamax(n, du, 1) == 12345
        want = oracle.sdot(n, u, 1, v, 1)
        assert abs(float(blas.sdot(n, du, 1, dv, 1)) - float(want)) <= 2 * n * E * float(np.sum(np.abs(u.astype(np.float64) * v)))
        assert abs(float(blas.snrm2(n, du, 1)) - float(oracle.snrm2(n, u, 1))) <= 2 * n * E * float(oracle.snrm2(n, u, 1))
        y = dv.clone()
        blas.inplace.saxpy_(n, 0.5, du, 1, y, 1)
        P.assert_same(y.to_host(), oracle.saxpy(n, 0.5, u, 1, v, 1))
        x2, y2 = du.clone(), dv.clone()
        blas.inplace.sswap_(n, x2, 1, y2, 1)
        P.assert_same(x2.to_host(), v); P.assert_same(y2.to_host(), u)
        blas.inplace.sscal_(n, 3.0, x2, 1)
        P.assert_same(x2.to_host(), oracle.sscal(n, 3.0, v, 1))
        x3, y3 = du.clone(), dv.clone()
        blas.inplace.srot_(n, x3, 1, y3, 1, np.cos(0.3), np.sin(0.3))
        wx, wy = oracle.srot(n, u, 1, v, 1, np.cos(0.3), np.sin(0.3))
        P.assert_same(x3.to_host(), wx); P.assert_same(y3.to_host(), wy)
        y4 = dv.clone()
        blas.inplace.scopy_(n - 3, du.view(3, n - 3), 1, y4.view(1, n - 3), 1)      # differently misaligned operands
        P.assert_same(y4.to_host(), np.concatenate([v[:1], u[3:], v[n - 2:]]))
    finally:
        ctx.set_launch("all", -1, 0, 0)


def test_pdl_overlap_is_invisible(blas, ctx, oracle):
    """Programmatic dependent launch lets the kernel after a reduction start under its tail: every result
    must be bit-identical to the serialized order, including when the next kernel overwrites the vector
    the reduction just read."""
    from lambda_blas_b200 import inplace as ip
    n = (1 << 22) + 5
    rng = np.random.default_rng(77)
    hx, hy = gen_unit(rng, n), gen_unit(rng, n)
    hx[0] = np.nan                                  # isamax must still answer 0 (NaN at index 0 stays)
    outs = {}
    for pdl in (False, True):
        ctx.set_pdl(pdl)
        x, y, z, r = ctx.to_device(hx), ctx.to_device(hy), ctx.alloc(n).fill(0.0), ctx.alloc(16).fill(0.0)
        before = ctx.pdl_count()
        for it in range(6):
            ip.sdot_async(n - 1, x.view(1, n - 1), 1, y.view(1, n - 1), 1, r.view(0, 2))
            ip.saxpy_(n, 0.25, x, 1, y, 1)                   # writes y right after a reduction read it
            ip.snrm2_async(n - 1, y.view(1, n - 1), 1, r.view(2, 2))
            ip.isamax_async(n, x, 1, r.view(4, 2))
            ip.sscal_(n, 1.5, x, 1)                           # overwrites x[0] the isamax tail depends on
            ip.sasum_async(n - 1, x.view(1, n - 1), 1, r.view(6, 2))
            ip.sdot_async(n - 1, x.view(1, n - 1), 1, x.view(1, n - 1), 1, r.view(8, 2))
            # every kernel releases its dependents at once, so each hazard class must be caught on the host:
            ip.scopy_(n, x, 1, z, 1)                          # independent of the sdot above: overlaps it
            ip.scopy_(n, y, 1, z, 1)                          # write after write (z): stores wait for the first copy
            ip.snrm2_async(n, z, 1, r.view(10, 2))            # read after write (z): ordinary launch
            ip.sswap_(n, z, 1, y, 1)                          # writes z and y right after a reduction read z
            ip.saxpy_(n, -0.5, y, 1, z, 1)                    # reads y, z the swap is still writing
            ip.sasum_async(n, z, 1, r.view(12, 2))
        outs[pdl] = (r.to_host(), x.to_host(), y.to_host(), z.to_host(), ctx.pdl_count() - before)
    ctx.set_pdl(True)
    assert outs[False][4] == 0 and outs[True][4] >= 6 * 6
    for a, b in zip(outs[False][:4], outs[True][:4]):
        P.assert_same(a, b)
    assert outs[True][0][4:6].view(np.int64)[0] == 0


def test_cpp_mirror_example(tmp_path):
    """The compiled host layer (cpp/lambda_blas.hpp over the C ABI) runs the reference's README example."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "lb_example")
    lib = os.path.join(root, "lambda_blas_b200")
    subprocess.run(["g++", "-std=c++17", os.path.join(root, "cpp", "example.cpp"), "-o", exe, "-L" + lib, "-llbk",
                    "-Wl,-rpath," + lib], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sdot = 32" in r.stdout and r.stdout.strip().endswith("OK")


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 9, 64, 1001, 4100, (1 << 20) + 1, (1 << 20) + 2])
@pytest.mark.parametrize("offs", [(0, 0), (1, 0), (0, 3), (2, 1), (3, 3)])
@pytest.mark.parametrize("dtype", P.DTYPES)
def test_reversed_pairing_paths(blas, ctx, oracle, n, offs, dtype):
    """(incx,incy) = (1,-1) and (-1,1): the reversed-lane vector path and its scalar fallback, all alignments."""
    ox, oy = offs
    rng = np.random.default_rng(n + 7 * ox + 31 * oy)
    E = float(P.eps_of(dtype))
    u, v = gen_unit(rng, n + ox, dtype), gen_unit(rng, n + oy, dtype)
    du, dv = ctx.to_device(u).view(ox, n), ctx.to_device(v).view(oy, n)
    hu, hv = u[ox:], v[oy:]
    for incx, incy in ((1, -1), (-1, 1)):
        want = oracle.sdot(n, hu, incx, hv, incy)
        assert abs(float(blas.sdot(n, du, incx, dv, incy)) - float(want)) <= 2 * n * E * float(np.sum(np.abs(hu.astype(np.float64) * hv[::-1])))
        if dtype == np.float32:
            assert abs(float(blas.sdsdot(n, 0.5, du, incx, dv, incy)) - float(oracle.sdsdot(n, 0.5, hu, incx, hv, incy))) <= float(np.spacing(np.float32(abs(oracle.sdsdot(n, 0.5, hu, incx, hv, incy)))))
        P.assert_same(blas.saxpy(n, 0.75, du, incx, dv, incy).to_host(), oracle.saxpy(n, 0.75, hu, incx, hv, incy))
        P.assert_same(blas.scopy(n, du, incx, dv, incy).to_host(), oracle.scopy(n, hu, incx, hv, incy))
        for got, wnt in zip(blas.sswap(n, du, incx, dv, incy), oracle.sswap(n, hu, incx, hv, incy)):
            P.assert_same(got.to_host(), wnt)
        for got, wnt in zip(blas.srot(n, du, incx, dv, incy, 0.6, 0.8), oracle.srot(n, hu, incx, hv, incy, 0.6, 0.8)):
            P.assert_same(got.to_host(), wnt)
        for args in ((1.5, 0.75, 2.0, 0.5), (1.0, 2.0, 0.5, 3.0), (1e-9, 1.0, 1.0, 1.0)):
            pre = "s" if dtype == np.float32 else "d"
            fl, ofl = getattr(blas, pre + "rotmg")(*args), getattr(oracle, pre + "rotmg")(*args)
            for got, wnt in zip(blas.srotm(fl, n, du, incx, dv, incy), oracle.srotm(ofl, n, hu, incx, hv, incy)):
                P.assert_same(got.to_host(), wnt)


@pytest.mark.parametrize("n", [1, 7, 8, 33, 1000, 65537])
@pytest.mark.parametrize("incs", [(1, 1), (-1, -1), (1, -1), (-1, 1), (2, 3), (-3, 1), (0, 1), (1, 0)])
@pytest.mark.parametrize("dtype", P.DTYPES)
def test_guard_bands_stay_untouched(blas, ctx, oracle, n, incs, dtype):
    """compute-sanitizer is closed on this pool, so out-of-bounds WRITES are looked for directly: every in-place
    routine works on a window in the middle of a larger buffer of sentinels, and the sentinels must survive."""
    from lambda_blas_b200 import inplace as ip
    incx, incy = incs
    G = 37                                              # odd guard: the window is misaligned on purpose
    lx, ly = 1 + (n - 1) * abs(incx), 1 + (n - 1) * abs(incy)
    rng = np.random.default_rng(n * 131 + incx * 7 + incy)
    sent = dtype(-12345.678)

    def fresh():
        bx, by = np.full(lx + 2 * G, sent, dtype), np.full(ly + 2 * G, sent, dtype)
        bx[G:G + lx] = gen_unit(rng, lx, dtype); by[G:G + ly] = gen_unit(rng, ly, dtype)
        dx, dy = ctx.to_device(bx), ctx.to_device(by)
        return bx, by, dx, dy, dx.view(G, lx), dy.view(G, ly)

    pre = "s" if dtype == np.float32 else "d"
    fl, ofl = getattr(blas, pre + "rotmg")(1.0, 2.0, 0.5, 3.0), getattr(oracle, pre + "rotmg")(1.0, 2.0, 0.5, 3.0)
    ops = [
        ("saxpy", lambda x, y: ip.saxpy_(n, 0.5, x, incx, y, incy), lambda hx, hy: (hx, oracle.saxpy(n, 0.5, hx, incx, hy, incy))),
        ("scopy", lambda x, y: ip.scopy_(n, x, incx, y, incy), lambda hx, hy: (hx, oracle.scopy(n, hx, incx, hy, incy))),
        ("sswap", lambda x, y: ip.sswap_(n, x, incx, y, incy), lambda hx, hy: oracle.sswap(n, hx, incx, hy, incy)),
        ("srot", lambda x, y: ip.srot_(n, x, incx, y, incy, 0.6, 0.8), lambda hx, hy: oracle.srot(n, hx, incx, hy, incy, 0.6, 0.8)),
        ("srotm", lambda x, y: ip.srotm_(fl, n, x, incx, y, incy), lambda hx, hy: oracle.srotm(ofl, n, hx, incx, hy, incy)),
    ]
    if incx > 0:
        ops.append(("sscal", lambda x, y: ip.sscal_(n, 1.5, x, incx), lambda hx, hy: (oracle.sscal(n, 1.5, hx, incx), hy)))
    for name, run, ref in ops:
        if (incx == 0 or incy == 0) and name in ("sswap", "srot", "srotm", "saxpy") and incy == 0 and name == "saxpy":
            pass
        bx, by, dx, dy, vx, vy = fresh()
        run(vx, vy)
        wx, wy = ref(bx[G:G + lx], by[G:G + ly])
        gx, gy = dx.to_host(), dy.to_host()
        assert np.all(gx[:G] == sent) and np.all(gx[G + lx:] == sent), (name, "x guard")
        assert np.all(gy[:G] == sent) and np.all(gy[G + ly:] == sent), (name, "y guard")
        P.assert_same(gx[G:G + lx], wx, name); P.assert_same(gy[G:G + ly], wy, name)


def test_sdsdot_async_matches_the_blocking_call(blas, ctx, oracle):
    """lbk_sdsdot_async (result left on the device) and lbk_sdsdot (Single.hs:231-262) are the same reduction."""
    from lambda_blas_b200 import inplace as ip
    rng = np.random.default_rng(5)
    n = 123457
    hx, hy = gen_unit(rng, n), gen_unit(rng, n)
    x, y, r = ctx.to_device(hx), ctx.to_device(hy), ctx.alloc(4)
    for incx, incy, m in ((1, 1, n), (2, -3, n // 3), (-1, 1, n)):
        ip.sdsdot_async(m, 0.625, x, incx, y, incy, r)
        got = r.to_host()[:1]
        P.assert_same(got, np.array([blas.sdsdot(m, 0.625, x, incx, y, incy)], dtype=np.float32))
        want = float(oracle.sdsdot(m, 0.625, hx, incx, hy, incy))
        assert abs(float(got[0]) - want) <= 2.0 ** -23 * abs(want) + 1e-30       # double accumulation: one float rounding


@pytest.mark.parametrize("fenced,threads,bps", [(0, 128, 8), (1, 256, 0), (1, 128, 100000)])
def test_general_stride_kernels_under_every_launch_shape(blas, ctx, oracle, fenced, threads, bps):
    """The general-stride map / reduce kernels give the oracle's bits whatever their launch shape (the fenced form of
    the map kernel is not a default: it made no difference on B200, profiles/r01_strided_launch_sweep.md)."""
    from lambda_blas_b200 import inplace as ip
    rng = np.random.default_rng(11)
    n = 40003
    ctx.set_launch("mapstrided", fenced, threads, bps)
    ctx.set_launch("redstrided", 0, threads, max(bps, 0) if bps < 1000 else 64)
    try:
        for incx, incy in ((2, 3), (-2, 5), (7, -1), (3, 3)):
            hx, hy = gen_unit(rng, 1 + (n - 1) * abs(incx)), gen_unit(rng, 1 + (n - 1) * abs(incy))
            x, y = ctx.to_device(hx), ctx.to_device(hy)
            ip.saxpy_(n, 0.75, x, incx, y, incy)
            want = oracle.saxpy(n, 0.75, hx, incx, hy, incy)
            P.assert_same(y.to_host(), want)
            ip.srot_(n, x, incx, y, incy, np.cos(0.3), np.sin(0.3))
            wx, wy = oracle.srot(n, hx, incx, want, incy, np.cos(0.3), np.sin(0.3))
            P.assert_same(x.to_host(), wx); P.assert_same(y.to_host(), wy)
            got = float(blas.sdot(n, x, incx, y, incy))
            ref = float(oracle.sdot(n, wx, incx, wy, incy))
            assert abs(got - ref) <= 2 * n * 2.0 ** -24 * float(np.sum(np.abs(P.sample(wx, n, incx) * P.sample(wy, n, incy))))
            if incx > 0:
                assert blas.isamax(n, x, incx) == oracle.isamax(n, wx, incx)
    finally:
        ctx.set_launch("mapstrided", -1, 0, 0)
        ctx.set_launch("redstrided", -1, 0, 0)


@pytest.mark.parametrize("dtype", P.DTYPES)
def test_walk_direction_is_invisible(blas, ctx, oracle, dtype):
    """lbk_set_traversal: elementwise kernels walk large vectors forwards, backwards, or from the end the previous kernel left in
    the L2 -- the results are the same bits, for every shape (aligned, reversed lanes, shifted loads, one element per access)"""
    from lambda_blas_b200 import inplace as ip
    n = (12 << 20) + 5                                  # 48 MiB of floats per operand: above the adaptive threshold
    rng = np.random.default_rng(17)
    hx, hy = gen_unit(rng, n + 4, dtype), gen_unit(rng, n + 4, dtype)
    fl = (blas.srotmg if dtype == np.float32 else blas.drotmg)(1.0, 2.0, 0.5, 3.0)
    outs = []
    for mode in (0, 1, 2):
        ctx.set_traversal(mode)
        try:
            x, y, res = ctx.to_device(hx), ctx.to_device(hy), ctx.alloc(4, dtype).fill(0.0)
            for _ in range(2):                          # twice: the adaptive mode alternates
                ip.sdot_async(n, x, 1, y, 1, res)       # a reduction in between: it always walks forwards
                ip.saxpy_(n, 0.5, x, 1, y, 1)
                ip.sscal_(n, 1.0009765625, x, 1)
                ip.srot_(n, x, 1, y, 1, 0.6, 0.8)
                ip.srotm_(fl, n, x, 1, y, 1)
                ip.sswap_(n, x, 1, y, 1)
                ip.saxpy_(n, 0.25, x, 1, y, -1)         # reversed lanes
                ip.saxpy_(n, 0.125, x.view(1, n), 1, y.view(0, n), 1)      # shifted loads
                ip.sswap_(n, x.view(3, n), 1, y.view(2, n), 1)             # one element per access
                ip.scopy_(n // 2, x, 1, y.view(n // 2, n // 2), 1)
            outs.append((x.to_host(), y.to_host(), res.to_host()))
        finally:
            ctx.set_traversal(1)
    for got in outs[1:]:
        for a, b in zip(got, outs[0]):
            P.assert_same(a, b)

"""Status codes of the C ABI: degenerate arguments are the reference's sentinels (status OK), everything the
reference would turn into undefined behaviour is a loud error -- never a silent wrong answer."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_sentinels_are_not_errors(blas, ctx):
    x = ctx.to_device([3.0, -4.0, 1.0])
    assert blas.sdot(0, x, 1, x, 1) == 0 and blas.sdot(-5, x, 1, x, 1) == 0          # empty fold
    assert blas.sdsdot(0, 2.5, x, 1, x, 1) == np.float32(2.5)                          # Single.hs:239
    assert blas.sasum(3, x, 0) == 0 and blas.sasum(3, x, -1) == 0                      # Single.hs:162
    assert blas.snrm2(0, x, 1) == 0 and blas.snrm2(3, x, -2) == 0 and blas.snrm2(1, x, 5) == 3.0   # Single.hs:181-182
    assert blas.isamax(0, x, 1) == -1 and blas.isamax(2, x, 0) == -1 and blas.isamax(1, x, 9) == 0  # Single.hs:218-219
    assert blas.isamax(3, x, 1) == 1
    # sscal with a non-positive increment: Single.hs:108-111
    np.testing.assert_array_equal(blas.sscal(3, 2.0, x, -1).to_host(), [3, -4, 1])      # n>1: unchanged
    np.testing.assert_array_equal(blas.sscal(1, 2.0, x, -1).to_host(), [6, -4, 1])      # n==1: x[0]*a
    np.testing.assert_array_equal(blas.sscal(3, 2.0, x, 0).to_host(), [6, -4, 1])
    np.testing.assert_array_equal(blas.saxpy(0, 2.0, x, 1, x, 1).to_host(), [3, -4, 1])
    e = ctx.to_device([])
    assert len(e) == 0 and blas.sdot(0, e, 1, e, 1) == 0 and blas.isamax(0, e, 1) == -1


def test_out_of_range_spans_are_refused(blas, ctx):
    x, y = ctx.to_device(np.arange(10)), ctx.to_device(np.arange(4))
    for call in (lambda: blas.sdot(5, x, 1, y, 1), lambda: blas.sdot(4, x, 4, y, 1), lambda: blas.sdot(4, x, -4, y, 1),
                 lambda: blas.saxpy(11, 1.0, x, 1, x.clone(), 1), lambda: blas.snrm2(6, x, 2), lambda: blas.isamax(3, y, 2),
                 lambda: blas.sswap(3, x, 1, y, 2), lambda: blas.srot(4, x, 1, y, -2, 1.0, 0.0)):
        with pytest.raises(blas.LbkError) as ei:
            call()
        assert ei.value.status == 4, ei.value                                           # LBK_ERR_RANGE
    assert blas.sdot(4, x, 3, y, -1) == np.float32(0 * 3 + 3 * 2 + 6 * 1 + 9 * 0)      # the largest legal span


def test_aliasing_in_place_is_refused_pure_is_fine(blas, ctx):
    from lambda_blas_b200 import inplace as ip
    x = ctx.to_device(np.arange(8, dtype=np.float32))
    with pytest.raises(blas.LbkError) as ei:
        ip.saxpy_(8, 1.0, x, 1, x, 1)
    assert ei.value.status == 5                                                         # LBK_ERR_ALIAS
    with pytest.raises(blas.LbkError):
        ip.sswap_(4, x.view(0, 6), 1, x.view(2, 6), 1)
    ip.sswap_(4, x.view(0, 4), 1, x.view(4, 4), 1)                                      # disjoint halves are fine
    np.testing.assert_array_equal(x.to_host(), [4, 5, 6, 7, 0, 1, 2, 3])
    # the pure forms take the same vector twice, as the reference does: saxpy n a u 1 u 1 = (1+a)*u
    np.testing.assert_array_equal(blas.saxpy(8, 2.0, x, 1, x, 1).to_host(), 3 * x.to_host())
    a, b = blas.sswap(8, x, 1, x, -1)
    np.testing.assert_array_equal(a.to_host(), x.to_host()[::-1])
    assert blas.sdot(8, x, 1, x, 1) == np.float32(140)                                  # reads may alias


def test_foreign_handles_and_bad_arguments(blas, ctx):
    other = blas.Context(0)
    try:
        x, z = ctx.to_device([1.0, 2.0]), other.to_device([1.0, 2.0])
        with pytest.raises(ValueError):
            blas.sdot(2, x, 1, z, 1)                       # host mirror: different contexts
        from lambda_blas_b200 import _ffi
        import ctypes as C
        out = C.c_float()
        st = _ffi.lib().lbk_sdot(ctx.handle, 2, x.handle, 1, z.handle, 1, C.byref(out))
        assert st == 3 and b"another context" in _ffi.lib().lbk_last_error(ctx.handle)   # LBK_ERR_ARG
        assert _ffi.lib().lbk_sdot(ctx.handle, 2, None, 1, x.handle, 1, C.byref(out)) == 3
        assert _ffi.lib().lbk_sdot(ctx.handle, 2, x.handle, 1, x.handle, 1, None) == 3
        with pytest.raises(blas.LbkError):
            ctx.alloc(-1)
        with pytest.raises(blas.LbkError):
            x.view(1, 5)
        with pytest.raises(blas.LbkError):
            ctx.set_launch("nosuchroutine", 0, 0, 0)
        with pytest.raises(blas.LbkError):
            ctx.set_launch("sdot", 99, 0, 0)
    finally:
        other.close()


def test_huge_counts_and_increments_do_not_wrap(blas, ctx):
    """1 + (n-1)*|inc| must not be computed in wrapping int64 arithmetic: a span that overflows is out of range, not tiny."""
    import ctypes as C
    from lambda_blas_b200 import _ffi
    L = _ffi.lib()
    x = ctx.to_device(np.arange(16, dtype=np.float32))
    out, iout = C.c_float(), C.c_int64()
    big = 1 << 62
    for n, inc in ((big, 4), (big + 1, 8), (3, big), (2, -(1 << 63)), (1 << 61, -8), ((1 << 63) - 1, (1 << 63) - 1)):
        assert L.lbk_sdot(ctx.handle, n, x.handle, inc, x.handle, 1, C.byref(out)) == 4, (n, inc)        # LBK_ERR_RANGE
        assert L.lbk_saxpy(ctx.handle, n, 1.0, x.handle, 1, x.handle, inc) == 4, (n, inc)
        if inc > 0:
            assert L.lbk_isamax(ctx.handle, n, x.handle, inc, C.byref(iout)) == 4, (n, inc)
    # n == 1 never applies the increment: any value is fine (Single.hs:91-96: first(1, inc) = 0)
    assert L.lbk_sdot(ctx.handle, 1, x.handle, (1 << 63) - 1, x.handle, -(1 << 63), C.byref(out)) == 0 and out.value == 0.0
    assert L.lbk_isamax(ctx.handle, 1, x.handle, 1 << 62, C.byref(iout)) == 0 and iout.value == 0
    v = C.c_void_p()
    assert L.lbk_vec_view(x.handle, (1 << 63) - 1, 2, C.byref(v)) == 3                                   # offset + len would wrap
    lo, hi, off = C.c_int64(), C.c_int64(), C.c_int64()
    assert L.lbk_shard_span(1 << 40, 8, 3, 1 << 62, 1 << 20, C.byref(lo), C.byref(hi), C.byref(off)) == 4


def test_pure_results_keep_the_shard_layout(blas):
    """A rank that holds a WHOLE short vector (global length below the world size) must still hand out sharded results:
    lbk_vec_alloc_like copies the layout instead of guessing it from the geometry."""
    m = blas.MultiContext([0, 0, 0])
    try:
        for n in (1, 2, 3, 7):
            x = m.to_device_sharded(np.arange(1, n + 1, dtype=np.float32))
            y = blas.sscal(n, 2.0, x, 1)
            assert y.is_sharded and len(y) == n
            np.testing.assert_array_equal(y.to_host(), 2 * np.arange(1, n + 1, dtype=np.float32))
            assert blas.sasum(n, y, 1) == n * (n + 1) and blas.isamax(n, y, 1) == n - 1
    finally:
        m.close()

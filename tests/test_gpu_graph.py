"""CUDA graphs (lbk_capture_begin / lbk_capture_end / lbk_graph_launch): a recorded chain of calls replays bit for bit
what the same calls do eagerly; what cannot be recorded is refused without spoiling the capture."""
import gc

import numpy as np
import pytest

import tests.test_gpu_parity as P
from tests.gen import gen_unit

pytestmark = pytest.mark.gpu


def chain(ip, n, x, y, z, res, flags):
    """a fixed sequence that touches every recordable kind of call"""
    ip.sscal_(n, 1.0009765625, x, 1)
    ip.saxpy_(n, 0.5, x, 1, y, 1)                       # reads what sscal wrote
    ip.scopy_(n, y, 1, z, 1)
    ip.srot_(n, x, 1, z, 1, 0.6, 0.8)
    ip.srotm_(flags, n, y, 1, z, 1)
    ip.sswap_(n - 3, x.view(3, n - 3), 1, y.view(1, n - 3), 1)      # differently misaligned views: the one-element shape
    ip.saxpy_(n - 3, 0.25, z.view(3, n - 3), 1, y.view(1, n - 3), 1)  # ... and the shifted-load one
    ip.saxpy_(n // 2, -0.5, x, 2, y, -1)                 # general stride
    ip.sdot_async(n, x, 1, y, 1, res)
    ip.snrm2_async(n, z, 1, res.view(2, 2))
    ip.isamax_async(n, y, 1, res.view(4, 2))


@pytest.mark.parametrize("n", [1000, 300007])
@pytest.mark.parametrize("pdl", [True, False])
@pytest.mark.parametrize("dtype", P.DTYPES)
def test_replay_equals_eager_bit_for_bit(blas, ctx, n, pdl, dtype):
    from lambda_blas_b200 import inplace as ip
    rng = np.random.default_rng(n)
    hx, hy, hz = gen_unit(rng, n, dtype), gen_unit(rng, n, dtype), gen_unit(rng, n, dtype)
    flags = (blas.srotmg if dtype == np.float32 else blas.drotmg)(1.5, 0.75, 2.0, 0.5)
    REPS = 3
    ctx.set_pdl(pdl)
    try:
        # eager: the chain REPS times
        x, y, z, res = ctx.to_device(hx), ctx.to_device(hy), ctx.to_device(hz), ctx.alloc(8, dtype).fill(0.0)
        for _ in range(REPS):
            chain(ip, n, x, y, z, res, flags)
        want = [v.to_host() for v in (x, y, z, res)]
        # recorded once, replayed REPS times on the same start values
        x, y, z, res = ctx.to_device(hx), ctx.to_device(hy), ctx.to_device(hz), ctx.alloc(8, dtype).fill(0.0)
        before = ctx.launch_count()
        with ctx.capture() as cap:
            chain(ip, n, x, y, z, res, flags)
        assert ctx.launch_count() == before, "recording must not run (or count) anything"
        P.assert_same(x.to_host(), hx)                   # nothing has run yet
        g = cap.graph
        assert g.kernels >= 11
        for _ in range(REPS):
            g.launch()
        assert ctx.launch_count() == before + REPS * g.kernels
        for got, w in zip((x, y, z, res), want):
            P.assert_same(got.to_host(), w)
    finally:
        ctx.set_pdl(True)


def test_what_cannot_be_recorded_is_refused_and_the_capture_survives(blas, ctx):
    from lambda_blas_b200 import inplace as ip
    n = 4096
    x, y, out = ctx.alloc(n).fill_hash(1), ctx.alloc(n).fill_hash(2), ctx.alloc(4)
    ev = ctx.event()
    host = np.zeros(n, np.float32)
    with ctx.capture() as cap:
        ip.saxpy_(n, 2.0, x, 1, y, 1)
        for bad in (lambda: blas.sdot(n, x, 1, y, 1),          # returns a scalar to the host
                    lambda: blas.saxpy(n, 2.0, x, 1, y, 1),    # pure: allocates its result
                    lambda: ctx.alloc(16),
                    lambda: x.clone(),
                    lambda: ctx.sync(),
                    lambda: ev.record(),
                    lambda: x.download(host),
                    lambda: blas.sasum(0, x, 1),               # even a sentinel result goes back to the host
                    lambda: ctx.capture().__enter__()):
            with pytest.raises(blas.LbkError) as e:
                bad()
            assert e.value.status == 6, str(e.value)
        ip.sdot_async(n, x, 1, y, 1, out)
        tmp = y.view(0, 16)
        del tmp
        gc.collect()                                        # a finaliser inside the capture
    hx, hy = blas.hash_fill_reference(n, 1), blas.hash_fill_reference(n, 2)
    cap.graph.launch()
    y1 = (hy + np.float32(2.0) * hx).astype(np.float32)
    P.assert_same(y.to_host(), y1)
    assert abs(float(out.to_host()[0]) - float(np.dot(hx.astype(np.float64), y1.astype(np.float64)))) < 1e-2
    with pytest.raises(blas.LbkError):
        blas._ffi.check(blas._ffi.lib().lbk_capture_end(ctx.handle, None), ctx.handle)      # not capturing any more


def test_owned_vectors_released_inside_a_capture_are_freed_afterwards(blas, ctx):
    from lambda_blas_b200 import inplace as ip
    n = 1 << 20
    x, y = ctx.alloc(n).fill_hash(1), ctx.alloc(n).fill_hash(2)
    doomed = ctx.alloc(n).fill_hash(3)
    with ctx.capture() as cap:
        ip.saxpy_(n, 1.0, x, 1, y, 1)
        del doomed
        gc.collect()
        ip.sscal_(n, 2.0, y, 1)
    cap.graph.launch().launch()
    hx, hy = blas.hash_fill_reference(n, 1), blas.hash_fill_reference(n, 2)
    want = ((hy + hx) * np.float32(2) + hx) * np.float32(2)
    P.assert_same(y.to_host(), want.astype(np.float32))
    again = ctx.alloc(n).fill_hash(4)                       # the pool still works
    P.assert_same(again.to_host(), blas.hash_fill_reference(n, 4))


def test_graph_outlives_context_close(blas):
    from lambda_blas_b200 import inplace as ip
    c = blas.Context(0)
    x = c.alloc(1024).fill(1.0)
    with c.capture() as cap:
        ip.sscal_(1024, 3.0, x, 1)
    cap.graph.launch()
    assert x.to_host()[5] == 3.0
    c.close()
    with pytest.raises((blas.LbkError, RuntimeError)):
        cap.graph.launch()                                   # the context is gone ...
    del cap, x                                               # ... and the handles are still safe to destroy
    gc.collect()


def test_multi_device_contexts_do_not_capture(blas):
    m = blas.MultiContext([0, 0])
    try:
        with pytest.raises(blas.LbkError) as e:
            m.capture().__enter__()
        assert e.value.status == 6
    finally:
        m.close()


def test_closing_a_context_in_the_middle_of_a_capture(blas):
    from lambda_blas_b200 import inplace as ip
    c = blas.Context(0)
    x = c.alloc(4096).fill(2.0)
    blas._ffi.check(blas._ffi.lib().lbk_capture_begin(c.handle), c.handle)
    ip.sscal_(4096, 3.0, x, 1)
    c.close()                                  # drops the recording; nothing ran
    c2 = blas.Context(0)                       # the device is fine afterwards
    y = c2.alloc(4096).fill(1.5)
    ip.sscal_(4096, 2.0, y, 1)
    assert y.to_host()[7] == 3.0
    c2.close()
    del x
    gc.collect()

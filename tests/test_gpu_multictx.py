"""The single-process multi-device context (lbk_init_multi): ONE host thread, vectors sharded over ranks, every routine of the
reference's call surface fanned out over per-device streams, reductions finished inside the kernels over peer mailboxes.

Listing device 0 twice (or three times) makes the ranks share a GPU, so the whole sharded path -- block partition, fan-out,
in-kernel record exchange, rank-ordered fold -- runs on the one-GPU box the driver tests on; with more GPUs the same tests run
over all of them as well.  Bars as everywhere: isamax and the elementwise routines bit-exact against the oracle on the WHOLE
vector, sums within 2*n*u*sum|.|, and every rank holds the same bits (test/Main.hs:37-42,59-79 are the properties)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 8))))
    return lists


@pytest.fixture(scope="module", params=["2x0", "3x0", "all"])
def mctx(request, blas):
    import torch
    n = torch.cuda.device_count()
    devs = {"2x0": [0, 0], "3x0": [0, 0, 0], "all": list(range(min(n, 8)))}[request.param]
    if request.param == "all" and n < 2:
        pytest.skip("one GPU: the distinct-device case needs gpurun --gpus 2")
    c = blas.MultiContext(devs)
    yield c
    c.close()


def bits(a):
    a = np.asarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def same(a, b):
    return np.array_equal(bits(a), bits(b))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 5, 1000, (1 << 20) + 3])
def test_sharded_routines_match_oracle(blas, oracle, mctx, n, dt):
    from lambda_blas_b200 import inplace as ip, sharded
    R = mctx.nranks
    u = float(np.finfo(dt).eps) / 2
    rng = np.random.default_rng(n)
    hx, hy = rng.uniform(-1, 1, n).astype(dt), rng.uniform(-1, 1, n).astype(dt)
    if n >= 1000:      # a maximum at the last element of shard 0, its tie at the first element of shard 1 and again at the very end
        edge = sharded.shard_range(n, R, 0)[1]
        hx[edge - 1], hx[edge], hx[n - 1] = 7.0, -7.0, 7.0
    x, y = mctx.to_device_sharded(hx), mctx.to_device_sharded(hy)
    assert x.is_sharded == (R > 1) or n == 0
    assert len(x) == n and x.global_len == n
    for r in range(R):      # the parts tile the vector in rank order
        lo, hi = sharded.shard_range(n, R, r)
        p = x.part(r)
        assert len(p) == hi - lo and p.shard_offset == lo and same(p.to_host(), hx[lo:hi])
    assert same(x.to_host(), hx)                                   # gather of one host buffer
    P = "s" if dt == np.float32 else "d"
    o = lambda name: getattr(oracle, P + name)  # noqa: E731
    prod = float(np.sum(np.abs(hx.astype(np.float64) * hy)))
    assert abs(float(blas.dot(n, x, 1, y, 1)) - float(o("dot")(n, hx, 1, hy, 1))) <= 2 * n * u * prod
    assert abs(float(blas.nrm2(n, x, 1)) - float(o("nrm2")(n, hx, 1))) <= 2 * n * u * float(o("nrm2")(n, hx, 1))
    assert abs(float(blas.asum(n, x, 1)) - float(o("asum")(n, hx, 1))) <= 2 * n * u * float(np.sum(np.abs(hx)))
    assert blas.iamax(n, x, 1) == getattr(oracle, "i" + P + "amax")(n, hx, 1)
    if dt == np.float32:
        want = oracle.sdsdot(n, 0.5, hx, 1, hy, 1)
        assert abs(float(blas.sdsdot(n, 0.5, x, 1, y, 1)) - float(want)) <= abs(float(want)) * 2.0 ** -23 + 1e-30
    if n >= 5:              # n smaller than the vector: the trailing ranks contribute empty records
        m = n - 2
        assert blas.iamax(m, x, 1) == getattr(oracle, "i" + P + "amax")(m, hx, 1)
        assert abs(float(blas.dot(m, x, 1, y, 1)) - float(o("dot")(m, hx, 1, hy, 1))) <= 2 * n * u * prod
    # the pure elementwise routines: new sharded vectors, bit-exact, arguments untouched
    assert same(blas.axpy(n, 1.000123, x, 1, y, 1).to_host(), o("axpy")(n, 1.000123, hx, 1, hy, 1))
    assert same(blas.scal(n, 3.0, x, 1).to_host(), o("scal")(n, 3.0, hx, 1))
    assert same(blas.copy(n, x, 1, y, 1).to_host(), o("copy")(n, hx, 1, hy, 1))
    a, b = blas.swap(n, x, 1, y, 1)
    assert same(a.to_host(), hy) and same(b.to_host(), hx)
    c_, s_ = dt(np.cos(0.3)), dt(np.sin(0.3))
    a, b = blas.rot(n, x, 1, y, 1, c_, s_)
    wa, wb = o("rot")(n, hx, 1, hy, 1, c_, s_)
    assert same(a.to_host(), wa) and same(b.to_host(), wb)
    for args in ((1.5, 0.75, 2.0, 0.5), (2.0, 0.5, 0.25, 3.0), (1e-9, 1.0, 1.0, 1.0)):      # FLAG0, FLAG1, FLAGNEG1 (rescaled)
        fl, ofl = getattr(blas, P + "rotmg")(*map(dt, args)), getattr(oracle, P + "rotmg")(*map(dt, args))
        a, b = blas.rotm(fl, n, x, 1, y, 1)
        wa, wb = o("rotm")(ofl, n, hx, 1, hy, 1)
        assert same(a.to_host(), wa) and same(b.to_host(), wb), fl
    assert same(x.to_host(), hx) and same(y.to_host(), hy)
    # in place, and equal positive increments (logical element i at global position i*inc: every pairing stays on one rank)
    for inc in (2, 3, 7):
        m = (n - 1) // inc + 1
        prod = float(np.sum(np.abs(hx[::inc][:m].astype(np.float64) * hy[::inc][:m])))
        assert abs(float(blas.dot(m, x, inc, y, inc)) - float(o("dot")(m, hx, inc, hy, inc))) <= 2 * m * u * prod + 1e-300
        assert blas.iamax(m, x, inc) == getattr(oracle, "i" + P + "amax")(m, hx, inc)
        y2 = y.clone()
        ip.saxpy_(m, -0.75, x, inc, y2, inc)
        assert same(y2.to_host(), o("axpy")(m, -0.75, hx, inc, hy, inc))
    # equal NEGATIVE increments pair the same elements as their absolute values: supported on shards, bit-exact for the maps
    for inc in (-1, -3):
        m = (n - 1) // abs(inc) + 1
        prod = float(np.sum(np.abs(hx[::abs(inc)][:m].astype(np.float64) * hy[::abs(inc)][:m])))
        assert abs(float(blas.dot(m, x, inc, y, inc)) - float(o("dot")(m, hx, inc, hy, inc))) <= 2 * m * u * prod + 1e-300
        assert same(blas.axpy(m, 0.5, x, inc, y, inc).to_host(), o("axpy")(m, 0.5, hx, inc, hy, inc))
        a, b = blas.rot(m, x, inc, y, inc, c_, s_)
        wa, wb = o("rot")(m, hx, inc, hy, inc, c_, s_)
        assert same(a.to_host(), wa) and same(b.to_host(), wb)
    # what would pair elements across devices is refused, not silently wrong
    if R > 1:
        for incx, incy in ((2, 3), (-2, -3), (1, -1), (0, 1)):
            with pytest.raises(blas.LbkError) as ei:
                blas.dot(max(n // 4, 1), x, incx, y, incy)
            assert ei.value.status == 6


def test_nan_rule_and_ties_across_ranks(blas, oracle, mctx):
    """isamax: a NaN wins only at logical index 0 (rank 0's first element); equal maxima on different ranks -> the lowest index."""
    from lambda_blas_b200 import sharded
    R, n = mctx.nranks, 4099
    rng = np.random.default_rng(7)
    base = rng.uniform(-1, 1, n).astype(np.float32)
    edges = [sharded.shard_range(n, R, r)[0] for r in range(R)]
    for name, planted in (("nan@0", {0: np.nan}), ("nan elsewhere", {edges[-1]: np.nan, 17: 3.0}),
                          ("tie on every rank", {e + 1: (-5.0 if i % 2 else 5.0) for i, e in enumerate(edges)}),
                          ("max at the last element", {n - 1: -9.0}), ("all equal", None)):
        h = base.copy() if planted is not None else np.full(n, 2.5, np.float32)
        for k, v in (planted or {}).items():
            h[k] = v
        x = mctx.to_device_sharded(h)
        assert blas.isamax(n, x, 1) == oracle.isamax(n, h, 1), name
        if not np.isnan(h).any():
            assert blas.snrm2(n, x, 1) == pytest.approx(float(oracle.snrm2(n, h, 1)), rel=1e-6)


def test_every_rank_holds_the_same_bits(blas, mctx):
    """Every rank folds the same records in the same order: each rank's own copy of the result (lbk_multi_result) is
    bit-identical to every other rank's, for all four reductions and both exchanges."""
    R, n = mctx.nranks, (1 << 18) + 5
    x, y = mctx.alloc_sharded(n).fill_hash(3), mctx.alloc_sharded(n).fill_hash(4)
    hx = blas.hash_fill_reference(n, 3)
    shared_gpu = len(set(mctx.devices)) < R
    for fused in ((False,) if shared_gpu else (True, False)):      # ranks that share a GPU never wait for one another in a kernel
        mctx.set_fused_exchange(fused)
        assert mctx.fused_exchange() == fused
        for nbytes, call in ((4, lambda: blas.sdot(n, x, 1, y, 1)), (4, lambda: blas.snrm2(n, x, 1)), (4, lambda: blas.sasum(n, x, 1)),
                             (8, lambda: blas.isamax(n, x, 1))):
            got = call()
            mine = [mctx.rank_result(r)[:nbytes] for r in range(R)]
            assert all(b == mine[0] for b in mine), mine
            want = np.frombuffer(mine[0], dtype=np.int64 if nbytes == 8 else np.float32)[0]
            assert want == got
        assert blas.isamax(n, x, 1) == int(np.argmax(np.abs(hx)))
    mctx.set_fused_exchange(not shared_gpu)


def test_ranks_driven_by_hand(blas, mctx):
    """The rank contexts (lbk_multi_rank) are one-device contexts whose sharded vectors are their blocks: a caller may drive them
    itself -- with the in-kernel exchange, i.e. on distinct GPUs; with the stream-ordered exchange the call is refused."""
    from lambda_blas_b200 import inplace as ip
    R, n = mctx.nranks, (1 << 18) + 5
    x, y = mctx.alloc_sharded(n).fill_hash(3), mctx.alloc_sharded(n).fill_hash(4)
    outs = [mctx.rank_context(r).alloc(8) for r in range(R)]
    if len(set(mctx.devices)) < R:
        with pytest.raises(blas.LbkError) as ei:
            ip.sdot_async(n, x.part(0), 1, y.part(0), 1, outs[0])
        assert ei.value.status == 6
        return
    for r in range(R):          # enqueue only: rank 0's kernel waits in its exchange until rank R-1's has been launched
        ip.sdot_async(n, x.part(r), 1, y.part(r), 1, outs[r])
    mctx.sync()
    words = [outs[r].to_host()[:1].view(np.uint32).tolist() for r in range(R)]
    assert all(w == words[0] for w in words), words
    assert np.array(words[0], dtype=np.uint32).view(np.float32)[0] == blas.sdot(n, x, 1, y, 1)


def test_unsharded_vectors_keep_full_semantics(blas, oracle, mctx):
    """to_device on a multi-device context = the whole vector on devices[0]: negative and zero increments work there."""
    rng = np.random.default_rng(11)
    n = 1001
    hx, hy = rng.uniform(-1, 1, 2 * n).astype(np.float32), rng.uniform(-1, 1, 3 * n).astype(np.float32)
    x, y = mctx.to_device(hx), mctx.to_device(hy)
    assert not x.is_sharded
    prod = float(np.sum(np.abs(hx[::2][:n].astype(np.float64) * hy[::3][:n][::-1])))
    assert abs(float(blas.sdot(n, x, 2, y, -3)) - float(oracle.sdot(n, hx, 2, hy, -3))) <= 2 * n * 2.0 ** -24 * prod
    assert same(blas.saxpy(n, 0.5, x, -2, y, 3).to_host(), oracle.saxpy(n, 0.5, hx, -2, hy, 3))
    assert same(blas.scopy(n, x, 1, y, 0).to_host(), oracle.scopy(n, hx, 1, hy, 0))
    assert blas.isamax(n, x.view(3, n), 1) == oracle.isamax(n, hx[3:3 + n], 1)
    if mctx.nranks > 1:
        with pytest.raises(blas.LbkError) as ei:
            blas.sdot(10, x, 1, mctx.to_device_sharded(hy), 1)          # sharded and unsharded do not mix
        assert ei.value.status == 6


def test_threaded_and_serial_fan_out_agree(blas, mctx):
    n = (1 << 20) + 1
    x, y = mctx.alloc_sharded(n).fill_hash(5), mctx.alloc_sharded(n).fill_hash(6)
    res = []
    for threads in (False, True, False):
        mctx.set_threads(threads)
        res.append((blas.sdot(n, x, 1, y, 1), blas.snrm2(n, x, 1), blas.isamax(n, x, 1), bits(blas.saxpy(n, 2.0, x, 1, y, 1).to_host()).sum()))
    assert res[0] == res[1] == res[2]


def test_a_missing_peer_times_out_instead_of_hanging(blas):
    """One rank makes a sharded call the other never makes: after the (shortened) watchdog the call fails with LBK_ERR_NCCL."""
    import time
    from lambda_blas_b200 import inplace as ip
    m = blas.MultiContext([0, 0], threads=False)
    try:
        m.set_fused_exchange(True)          # the in-kernel exchange, asked for explicitly: ONE kernel, waiting alone
        m.set_exchange_timeout(100)
        x = m.alloc_sharded(1 << 12).fill_hash(1)
        r0 = m.rank_context(0)
        out = r0.alloc(8)
        t0 = time.time()
        ip.sasum_async(1 << 12, x.part(0), 1, out)           # rank 1 never calls
        with pytest.raises(blas.LbkError) as ei:
            r0.sync()
        assert ei.value.status == 2 and 0.08 < time.time() - t0 < 5.0
    finally:
        m.close()


def test_handles_outlive_their_context(blas):
    """lbk_shutdown with live handles only closes the context; the last free releases it (GC finalisers run late)."""
    from lambda_blas_b200 import _ffi
    L = _ffi.lib()
    for make in (lambda: blas.Context(0), lambda: blas.MultiContext([0, 0])):
        c = make()
        v = c.alloc_sharded(1000).fill_hash(1)
        e = c.event().record()
        h = c._h
        c.close()                                               # lbk_shutdown: v and e are still alive
        out = C.c_float()
        assert L.lbk_sasum(h, 10, v.handle, 1, C.byref(out)) == 3      # LBK_ERR_ARG: the context has been shut down
        assert b"shut down" in L.lbk_last_error(h)
        w = C.c_void_p()
        assert L.lbk_vec_alloc(h, 10, C.byref(w)) == 3 and L.lbk_vec_clone(v.handle, C.byref(w)) == 3
        assert L.lbk_shutdown(h) == 0                                   # idempotent while handles are alive
        assert L.lbk_event_destroy(e._h) == 0
        e._h = None
        assert L.lbk_vec_free(v.handle) == 0                            # the last handle: the context goes with it
        v._h = None


def test_host_register_and_caller_device_is_restored(blas, ctx):
    import torch
    a = np.random.default_rng(0).uniform(-1, 1, 1 << 20).astype(np.float32)
    with blas.registered(a):
        v = ctx.alloc(a.size)
        v.upload_async(a)
        ctx.sync()
    assert same(v.to_host(), a)
    if torch.cuda.device_count() >= 2:
        torch.cuda.set_device(1)
        assert blas.snrm2(a.size, v, 1) > 0                     # runs on device 0 ...
        assert torch.cuda.current_device() == 1                 # ... and leaves the caller's device alone
        torch.cuda.set_device(0)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_staged_transfers_of_pageable_memory_round_trip(blas, ctx, dt):
    """lbk_vec_upload / lbk_vec_download of a large PAGEABLE buffer go through the library's pinned staging lanes (ragged sizes:
    a partial last chunk, fewer chunks than lanes); the bytes must come back unchanged, also through a sharded vector."""
    rng = np.random.default_rng(3)
    for n in ((16 << 20) // np.dtype(dt).itemsize, (40 << 20) // np.dtype(dt).itemsize + 12345, (8 << 20) // np.dtype(dt).itemsize - 7):
        h = rng.integers(0, 2 ** 31, n).astype(dt)
        v = ctx.alloc(n, dt)
        v.upload(h)
        assert same(v.to_host(), h)
        assert float(blas.asum(n, v, 1)) == pytest.approx(float(np.sum(h, dtype=np.float64)), rel=1e-5)
    m = blas.MultiContext([0, 0, 0])
    try:
        h = rng.integers(0, 2 ** 31, (100 << 20) // 4 + 3).astype(np.float32)
        assert same(m.to_device_sharded(h).to_host(), h)
    finally:
        m.close()

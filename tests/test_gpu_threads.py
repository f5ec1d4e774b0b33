"""One context per host thread (the threading rule of include/lbk.h): several threads, each with its own context on the same GPU,
run the whole call surface at once -- what a multi-threaded Haskell (or C++) host would do -- and every thread must get the answers
it gets alone.  Guards the library's process-wide state (context life cycle, per-instantiation caches, staging lanes, error strings)."""
import threading

import numpy as np
import pytest

import tests.test_gpu_parity as P
from tests.gen import gen_unit

pytestmark = pytest.mark.gpu


def workload(blas, oracle, seed, out):
    from lambda_blas_b200 import inplace as ip
    try:
        ctx = blas.Context(0)
        rng = np.random.default_rng(seed)
        n = int(rng.choice([1000, 65537, (1 << 20) + 3]))
        for rep in range(6):
            hx, hy = gen_unit(rng, n, np.float32), gen_unit(rng, n, np.float32)
            x, y = ctx.to_device(hx), ctx.to_device(hy)
            d = blas.sdot(n, x, 1, y, 1)
            assert abs(float(d) - float(oracle.sdot(n, hx, 1, hy, 1))) <= 2 * n * 2.0 ** -24 * float(np.sum(np.abs(hx.astype(np.float64) * hy)))
            assert blas.isamax(n, x, 1) == oracle.isamax(n, hx, 1)
            P.assert_same(blas.saxpy(n, 0.5, x, 1, y, 1).to_host(), oracle.saxpy(n, 0.5, hx, 1, hy, 1))
            m = 1 + (n - 1) // 3
            P.assert_same(blas.sscal(m, 1.5, x, 3).to_host(), oracle.sscal(m, 1.5, hx, 3))
            assert blas.isamax(m, x, 3) == oracle.isamax(m, hx, 3)
            with ctx.capture() as cap:
                ip.saxpy_(n, 0.25, x, 1, y, 1)
                ip.srot_(n, x, 1, y, 1, 0.6, 0.8)
            cap.graph.launch()
            wy = oracle.saxpy(n, 0.25, hx, 1, hy, 1)
            wx, wy = oracle.srot(n, hx, 1, wy, 1, 0.6, 0.8)
            P.assert_same(x.to_host(), wx); P.assert_same(y.to_host(), wy)
            with pytest.raises(blas.LbkError):
                blas.sdot(n + 1, x, 1, y, 1)                 # an error on this thread's context stays on this thread
        big = gen_unit(rng, (5 << 20) + 1, np.float32)       # 20 MiB: the staged (multi-lane) synchronous transfer
        P.assert_same(ctx.to_device(big).to_host(), big)
        ctx.close()
        out[seed] = "ok"
    except BaseException as e:      # noqa: BLE001
        out[seed] = repr(e)


def test_contexts_on_concurrent_host_threads(blas, oracle):
    out = {}
    threads = [threading.Thread(target=workload, args=(blas, oracle, s, out)) for s in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert out == {s: "ok" for s in range(6)}, out

"""Elements drawn from ALL bit patterns (NaN payloads, infinities, denormals, signed zeros): the elementwise routines and iamax must
still match the oracle bit for bit (any NaN matches any NaN) -- no flush-to-zero, no fast-math shortcuts, no value-dependent paths
that the reference's generators never reach.  Unit stride (aligned, shifted, reversed lanes) and general stride, both precisions."""
import numpy as np
import pytest

import tests.test_gpu_parity as P
from tests.gen import span

pytestmark = pytest.mark.gpu


def patterns(rng, n, dt):
    U = np.uint32 if dt == np.float32 else np.uint64
    kind = int(rng.integers(3))
    if kind == 0:
        return rng.integers(0, 2 ** 32 if dt == np.float32 else 2 ** 63, n, dtype=np.uint64).astype(U).view(dt).copy()
    if kind == 1:
        with np.errstate(over="ignore"):
            return (rng.standard_normal(n) * 10.0 ** rng.integers(-40, 39, n)).astype(dt)
    return rng.choice(np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 1e-40, 3e38, 2.0 ** 60, 2.0 ** -70]), n).astype(dt)


@pytest.mark.parametrize("dtype", P.DTYPES)
@pytest.mark.parametrize("seed", range(4))
def test_elementwise_and_iamax_on_arbitrary_bit_patterns(blas, ctx, oracle, seed, dtype):
    rng = np.random.default_rng(4000 + seed)
    pre = "s" if dtype == np.float32 else "d"
    for it in range(40):
        n = int(rng.choice([1, 3, 17, 100, 1000, 4099, 70001]))
        ix, iy = (1, 1) if it % 3 == 0 else (int(rng.integers(-3, 4)), int(rng.integers(-3, 4)))
        ox, oy = (int(rng.integers(0, 4)), int(rng.integers(0, 4))) if it % 2 else (0, 0)
        hx, hy = patterns(rng, span(n, ix, 1), dtype), patterns(rng, span(n, iy, 1), dtype)
        bx, by = np.zeros(hx.size + ox, dtype), np.zeros(hy.size + oy, dtype)
        bx[ox:], by[oy:] = hx, hy
        x, y = ctx.to_device(bx).view(ox, hx.size), ctx.to_device(by).view(oy, hy.size)
        a = patterns(rng, 1, dtype)[0]
        c, s = patterns(rng, 2, dtype)
        tag = (n, ix, iy, ox, oy)
        with np.errstate(all="ignore"):
            P.assert_same(blas.saxpy(n, a, x, ix, y, iy).to_host(), oracle.saxpy(n, a, hx, ix, hy, iy), ("saxpy",) + tag)
            P.assert_same(blas.scopy(n, x, ix, y, iy).to_host(), oracle.scopy(n, hx, ix, hy, iy), ("scopy",) + tag)
            P.assert_same(blas.sscal(n, a, x, ix).to_host(), oracle.sscal(n, a, hx, ix), ("sscal",) + tag)
            if ix != 0 and iy != 0:
                for got, want in zip(blas.sswap(n, x, ix, y, iy), oracle.sswap(n, hx, ix, hy, iy)):
                    P.assert_same(got.to_host(), want, ("sswap",) + tag)
                for got, want in zip(blas.srot(n, x, ix, y, iy, c, s), oracle.srot(n, hx, ix, hy, iy, c, s)):
                    P.assert_same(got.to_host(), want, ("srot",) + tag)
                fl = getattr(blas, pre + "rotmg")(1.5, 0.75, 2.0, 0.5 + it % 3)
                ofl = getattr(oracle, pre + "rotmg")(1.5, 0.75, 2.0, 0.5 + it % 3)
                for got, want in zip(blas.srotm(fl, n, x, ix, y, iy), oracle.srotm(ofl, n, hx, ix, hy, iy)):
                    P.assert_same(got.to_host(), want, ("srotm",) + tag)
            assert blas.isamax(n, x, ix) == oracle.isamax(n, hx, ix), ("isamax",) + tag

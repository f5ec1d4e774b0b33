"""Unit-stride operands misaligned DIFFERENTLY, read-only x the odd one out: aligned 16-byte packs of x + a lane
shift (lbk_kernels.cuh, shift_pack) for dot / sdsdot / axpy / copy / out-of-place scal.  Every shift (1..3 elements
for Float, 1 for Double), sizes around the tile and warp edges, both unrolls, guard bands around the in-place
destinations, and bit-identity with the one-element path the shifted one replaces."""
import numpy as np
import pytest

import tests.test_gpu_parity as P
from tests.gen import gen_unit

pytestmark = pytest.mark.gpu

SENT = -12345.678


def windows(ctx, rng, n, ox, oy, dtype):
    """x and y as views at element offsets ox / oy inside sentinel-filled buffers (allocations are 256-byte aligned)"""
    G = 40
    bx, by = np.full(n + 2 * G, SENT, dtype), np.full(n + 2 * G, SENT, dtype)
    bx[ox:ox + n], by[oy:oy + n] = gen_unit(rng, n, dtype), gen_unit(rng, n, dtype)
    dx, dy = ctx.to_device(bx), ctx.to_device(by)
    return bx, by, dx, dy, dx.view(ox, n), dy.view(oy, n)


@pytest.mark.parametrize("n", [64, 65, 127, 1000, 2048, 2051, 4099, 65537, (1 << 20) + 3])
@pytest.mark.parametrize("offs", [(1, 0), (2, 0), (3, 0), (0, 1), (5, 2), (7, 4), (2, 3), (1, 8)])
@pytest.mark.parametrize("dtype", P.DTYPES)
def test_shifted_loads_match_the_oracle(blas, ctx, oracle, n, offs, dtype):
    from lambda_blas_b200 import inplace as ip
    ox, oy = offs
    vec = 4 if dtype == np.float32 else 2
    differently = (ox - oy) % vec != 0
    rng = np.random.default_rng(n * 64 + ox * 8 + oy)
    E = float(P.eps_of(dtype))
    for unroll_map, unroll_red in ((19, 0), (20, 1)):            # 16-byte packs x 4 / x 2
        ctx.set_launch("mapshifted", unroll_map, 0, 0); ctx.set_launch("redshifted", unroll_red, 0, 0)
        try:
            bx, by, dx, dy, vx, vy = windows(ctx, rng, n, ox, oy, dtype)
            hx, hy = bx[ox:ox + n], by[oy:oy + n]
            before = ctx.shifted_count()
            got = blas.sdot(n, vx, 1, vy, 1)
            assert abs(float(got) - float(oracle.sdot(n, hx, 1, hy, 1))) <= 2 * n * E * float(np.sum(np.abs(hx.astype(np.float64) * hy)))
            if dtype == np.float32:
                w = float(oracle.sdsdot(n, 0.25, hx, 1, hy, 1))
                assert abs(float(blas.sdsdot(n, 0.25, vx, 1, vy, 1)) - w) <= float(np.spacing(np.float32(abs(w))))
            P.assert_same(blas.saxpy(n, 1.000123, vx, 1, vy, 1).to_host(), oracle.saxpy(n, 1.000123, hx, 1, hy, 1))     # pure: y misaligned, result aligned
            P.assert_same(blas.sscal(n, 1.5, vx, 1).to_host(), oracle.sscal(n, 1.5, hx, 1))                              # pure: x misaligned, result aligned
            P.assert_same(blas.scopy(n, vx, 1, vy, 1).to_host(), oracle.scopy(n, hx, 1, hy, 1))
            ip.saxpy_(n, 0.5, vx, 1, vy, 1)
            gy = dy.to_host()
            assert np.all(gy[:oy] == dtype(SENT)) and np.all(gy[oy + n:] == dtype(SENT)), "saxpy wrote outside y"
            P.assert_same(gy[oy:oy + n], oracle.saxpy(n, 0.5, hx, 1, hy, 1))
            ip.scopy_(n, vx, 1, vy, 1)
            gy = dy.to_host()
            assert np.all(gy[:oy] == dtype(SENT)) and np.all(gy[oy + n:] == dtype(SENT)), "scopy wrote outside y"
            P.assert_same(gy[oy:oy + n], hx)
            P.assert_same(dx.to_host(), bx)                                                                              # x is read-only
            used = ctx.shifted_count() - before
            if differently:      # at least sdot, sdsdot and the two in-place calls (the pure forms also depend on the result's alignment)
                assert used >= (4 if dtype == np.float32 else 3), f"the shifted path was not taken ({used} launches)"
        finally:
            ctx.set_launch("mapshifted", -1, 0, 0); ctx.set_launch("redshifted", -1, 0, 0)


@pytest.mark.parametrize("dtype", P.DTYPES)
def test_shifted_and_one_element_paths_agree_bit_for_bit(blas, ctx, dtype):
    """elementwise results do not depend on the path; turning the shifted path off must not change a bit"""
    from lambda_blas_b200 import inplace as ip
    n, ox, oy = 300007, 3, 1 if dtype == np.float32 else 0
    rng = np.random.default_rng(11)
    outs = []
    for variant in (-1, 8):                   # default shape / a one-element variant = path off
        ctx.set_launch("mapshifted", variant, 0, 0)
        try:
            r = np.random.default_rng(11)
            bx, by, dx, dy, vx, vy = windows(ctx, r, n, ox, oy, dtype)
            before = ctx.shifted_count()
            ip.saxpy_(n, 0.75, vx, 1, vy, 1)
            a = dy.to_host()
            ip.scopy_(n, vx, 1, vy, 1)
            outs.append((a, dy.to_host(), blas.sscal(n, 3.0, vx, 1).to_host(), ctx.shifted_count() - before))
        finally:
            ctx.set_launch("mapshifted", -1, 0, 0)
    assert outs[0][3] == 3 and outs[1][3] == 0
    for got, want in zip(outs[0][:3], outs[1][:3]):
        P.assert_same(got, want)
    del rng


def test_small_or_doubly_written_operands_keep_the_one_element_path(blas, ctx, oracle):
    """n below a few packs, and routines that write both vectors (swap, rot, rotm), are not shifted"""
    from lambda_blas_b200 import inplace as ip
    rng = np.random.default_rng(3)
    for n in (5, 30, 100000):
        bx, by, dx, dy, vx, vy = windows(ctx, rng, n, 1, 0, np.float32)
        before = ctx.shifted_count()
        ip.sswap_(n, vx, 1, vy, 1)
        ip.srot_(n, vx, 1, vy, 1, 0.6, 0.8)
        assert ctx.shifted_count() == before
        wx, wy = oracle.srot(n, by[0:n], 1, bx[1:1 + n], 1, 0.6, 0.8)
        P.assert_same(dx.to_host()[1:1 + n], wx); P.assert_same(dy.to_host()[0:n], wy)
        if n < 64:
            blas.sdot(n, vx, 1, vy, 1)
            assert ctx.shifted_count() == before

"""Random call sequences with programmatic dependent launch on and off must leave identical bits.

Every unit-stride kernel releases its dependents as its first instruction, so which kernels may overlap is
decided on the host (lbk_lib.cu: pdl_decide, hazard groups, TriggerPredictor).  A hole there shows up as a
kernel reading data a predecessor has not written yet, or overwriting data a predecessor is still reading --
i.e. as a difference from the serialised order.  The sequences mix reductions and updates over a few vectors
and overlapping windows of them, so that true, anti and output dependencies, chains and independent pairs all
occur, at sizes where a kernel is short enough for several to be in flight.
"""
import numpy as np
import pytest

from tests.gen import gen_unit

pytestmark = pytest.mark.gpu


def run_sequence(blas, ctx, seed, pdl, dtype):
    from lambda_blas_b200 import inplace as ip
    rng = np.random.default_rng(seed)
    n = int(rng.choice([1 << 16, (1 << 18) + 24, (1 << 20) + 8, 1 << 21]))
    ctx.set_pdl(pdl)
    vecs = [ctx.to_device(gen_unit(np.random.default_rng(1000 * seed + k), n, dtype)) for k in range(4)]
    res = ctx.alloc(64, dtype).fill(0.0)
    before = ctx.pdl_count()

    def window():
        """A whole vector or a window of one (offsets keep 32-byte alignment so the wide kernels run)."""
        v = vecs[int(rng.integers(len(vecs)))]
        if rng.random() < 0.5:
            return v, n
        off = 8 * int(rng.integers(0, n // 16))
        m = int(rng.integers(n // 4, n - off + 1))
        return v.view(off, m), m

    def two():
        while True:
            (a, na), (b, nb) = window(), window()
            m = min(na, nb)
            lo_a, lo_b = a.ptr, b.ptr
            esz = np.dtype(dtype).itemsize
            if lo_a + m * esz <= lo_b or lo_b + m * esz <= lo_a:       # in-place two-vector updates must not alias
                return a, b, m

    def strided(inc):
        """a window and the number of elements an increment of `inc` can visit in it"""
        a, m = window()
        return a, 1 + (m - 1) // abs(inc)

    for step in range(48):
        op = int(rng.integers(18))
        slot = res.view(2 * int(rng.integers(30)), 2)
        if op == 0:
            a, b, m = two(); ip.sdot_async(m, a, 1, b, 1, slot)
        elif op == 1:
            a, m = window(); ip.snrm2_async(m, a, 1, slot)
        elif op == 2:
            a, m = window(); ip.sasum_async(m, a, 1, slot)
        elif op == 3:
            a, m = window(); ip.isamax_async(m, a, 1, slot)
        elif op == 4:
            a, b, m = two(); ip.saxpy_(m, 0.5 - rng.random(), a, 1, b, 1)
        elif op == 5:
            a, m = window(); ip.sscal_(m, 1.0 + 0.25 * (rng.random() - 0.5), a, 1)
        elif op == 6:
            a, b, m = two(); ip.scopy_(m, a, 1, b, 1)
        elif op == 7:
            a, b, m = two(); ip.sswap_(m, a, 1, b, 1)
        elif op == 8:
            a, b, m = two(); ip.srot_(m, a, 1, b, 1, np.cos(0.3), np.sin(0.3))
        elif op == 9:
            a, b, m = two(); ip.saxpy_(m, 0.25, a, 1, b, -1)             # the reversed-pair kernel
        elif op == 10:
            a, m = window(); ip.sdot_async(m, a, 1, a, 1, slot)          # both operands the same vector
        elif op == 11:                                                  # the pure forms: a NEW vector per call, the old one is
            i, j = (int(v) for v in rng.choice(len(vecs), 2, replace=False))    # released in stream order and its block reused
            vecs[j] = blas.saxpy(n, 0.125, vecs[i], 1, vecs[j], 1)
        elif op == 12:
            j = int(rng.integers(len(vecs)))
            vecs[j] = blas.sscal(n - 5, 0.75, vecs[j], 1)
        elif op == 13:                                                  # the general-stride kernels are members of the chain too
            inc = int(rng.choice([2, 3, -2, 7]))
            a, m = strided(inc); ip.sscal_(m, 1.0 + 0.125 * (rng.random() - 0.5), a, abs(inc))
        elif op == 14:
            a, b, m = two(); inc = int(rng.choice([2, -3, 5]))
            k = 1 + (m - 1) // abs(inc)
            ip.saxpy_(k, 0.25 - rng.random() / 2, a, inc, b, int(rng.choice([1, -1, 2])) if k * 2 <= m else 1)
        elif op == 15:
            a, b, m = two(); inc = int(rng.choice([2, 3, -2]))
            ip.sdot_async(1 + (m - 1) // abs(inc), a, inc, b, inc, slot)
        elif op == 16:
            inc = int(rng.choice([2, 5]))
            a, m = strided(inc); ip.isamax_async(m, a, inc, slot)
        else:
            inc = int(rng.choice([2, 4]))
            a, m = strided(inc); ip.snrm2_async(m, a, inc, slot)
    out = [v.to_host() for v in vecs] + [res.to_host()]
    return out, ctx.pdl_count() - before


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("seed", range(8))
def test_random_sequences_are_order_invariant(blas, ctx, seed, dtype):
    try:
        serial, zero = run_sequence(blas, ctx, seed, False, dtype)
        overlapped, launches = run_sequence(blas, ctx, seed, True, dtype)
    finally:
        ctx.set_pdl(True)
    assert zero == 0 and launches > 0
    U = np.uint32 if dtype == np.float32 else np.uint64
    for k, (a, b) in enumerate(zip(serial, overlapped)):
        bad = np.flatnonzero(a.view(U) != b.view(U))
        assert bad.size == 0, f"seed {seed}: buffer {k} differs at {bad[:8]} (of {bad.size})"

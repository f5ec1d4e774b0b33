"""tests/golden/ref_golden_mid.json -- the reference's own source evaluated (oracle/hs_eval.py) at n = 4099 / 20011, where the CUDA
path runs its multi-block kernels.  Inputs come from the counter hash (seed 1 / 2), vector results are compared through SHA-256 of
their bytes (bit for bit), reductions as bit patterns (CPU oracle: equal; CUDA path: within the reduction bound, sdsdot 1 ulp)."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests.gen import counter_hash, span

HERE = os.path.dirname(os.path.abspath(__file__))
G = json.load(open(os.path.join(HERE, "golden", "ref_golden_mid.json")))
CASES = [pytest.param(c, id=f"{c['dtype']}-{c['n']}-{c['incx']}-{c['incy']}") for c in G["cases"]]
MGR_FLAG = {"FLAGNEG2": -2, "FLAGNEG1": -1, "FLAG0": 0, "FLAG1": 1}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def val(bits, dt):
    return np.array([bits], dtype=np.uint32 if dt == np.float32 else np.uint64).view(dt)[0]


def inputs(c):
    dt = np.float32 if c["dtype"] == "float32" else np.float64
    n, ix, iy = c["n"], c["incx"], c["incy"]
    return dt, n, ix, iy, counter_hash(span(n, ix, 2), 1, dtype=dt), counter_hash(span(n, iy, 1), 2, dtype=dt)


def test_the_generator_of_the_inputs_is_the_device_generator(blas):
    for dt in (np.float32, np.float64):
        for seed in (1, 2):
            assert np.array_equal(counter_hash(5000, seed, dtype=dt), blas.hash_fill_reference(5000, seed, dtype=dt))


@pytest.mark.parametrize("c", CASES)
def test_oracle_reproduces_the_reference_at_mid_sizes(oracle, c):
    dt, n, ix, iy, x, y = inputs(c)
    a, cs, sn = dt(G["a"]), dt(G["c"]), dt(G["s"])
    assert val(c["sdot"], dt).tobytes() == oracle.dot(n, x, ix, y, iy).tobytes()
    assert sha(oracle.copy(n, x, ix, y, iy)) == c["scopy"] and sha(oracle.axpy(n, a, x, ix, y, iy)) == c["saxpy"]
    if "sdsdot" in c:
        assert val(c["sdsdot"], dt).tobytes() == oracle.sdsdot(n, a, x, ix, y, iy).tobytes()
    if "sswap" in c:
        m = (oracle.srotmg if dt == np.float32 else oracle.drotmg)(*G["rotmg_args"])
        assert m.flag == MGR_FLAG[c["mgr"]]
        assert [sha(v) for v in oracle.swap(n, x, ix, y, iy)] == c["sswap"]
        assert [sha(v) for v in oracle.rot(n, x, ix, y, iy, cs, sn)] == c["srot"]
        assert [sha(v) for v in oracle.rotm(m, n, x, ix, y, iy)] == c["srotm"]
    if "sasum" in c:
        assert val(c["sasum"], dt).tobytes() == oracle.asum(n, x, ix).tobytes()
        assert val(c["snrm2"], dt).tobytes() == oracle.nrm2(n, x, ix).tobytes()
        assert oracle.iamax(n, x, ix) == c["isamax"] and sha(oracle.scal(n, a, x, ix)) == c["sscal"]


@pytest.mark.gpu
@pytest.mark.parametrize("offs", [(0, 0), (1, 0), (3, 2)], ids=["aligned", "x+1", "x+3,y+2"])
@pytest.mark.parametrize("c", CASES)
def test_cuda_path_matches_the_reference_at_mid_sizes(blas, ctx, c, offs):
    """the vectors sit at element offsets `offs` inside larger allocations: aligned packs, shifted loads, one-element shapes"""
    dt, n, ix, iy, x, y = inputs(c)
    E = 2.0 ** -24 if dt == np.float32 else 2.0 ** -53
    a, cs, sn = dt(G["a"]), dt(G["c"]), dt(G["s"])
    ox, oy = offs
    bx, by = np.zeros(x.size + ox, dt), np.zeros(y.size + oy, dt)
    bx[ox:], by[oy:] = x, y
    dx, dy = ctx.to_device(bx).view(ox, x.size), ctx.to_device(by).view(oy, y.size)
    fx = x[(0 if ix > 0 else (1 - n) * ix) + np.arange(n) * ix].astype(np.float64)
    fy = y[(0 if iy > 0 else (1 - n) * iy) + np.arange(n) * iy].astype(np.float64)
    assert abs(float(blas.sdot(n, dx, ix, dy, iy)) - float(val(c["sdot"], dt))) <= 2 * n * E * float(np.sum(np.abs(fx * fy)))
    assert sha(blas.scopy(n, dx, ix, dy, iy).to_host()) == c["scopy"]
    assert sha(blas.saxpy(n, a, dx, ix, dy, iy).to_host()) == c["saxpy"]
    if "sdsdot" in c:
        w = float(val(c["sdsdot"], dt))
        assert abs(float(blas.sdsdot(n, a, dx, ix, dy, iy)) - w) <= float(np.spacing(np.float32(abs(w))))
    if "sswap" in c:
        m = (blas.srotmg if dt == np.float32 else blas.drotmg)(*G["rotmg_args"])
        assert m.flag == MGR_FLAG[c["mgr"]]
        assert [sha(v.to_host()) for v in blas.sswap(n, dx, ix, dy, iy)] == c["sswap"]
        assert [sha(v.to_host()) for v in blas.srot(n, dx, ix, dy, iy, cs, sn)] == c["srot"]
        assert [sha(v.to_host()) for v in blas.srotm(m, n, dx, ix, dy, iy)] == c["srotm"]
    if "sasum" in c:
        assert abs(float(blas.sasum(n, dx, ix)) - float(val(c["sasum"], dt))) <= 2 * n * E * float(np.sum(np.abs(fx)))
        wn = float(val(c["snrm2"], dt))
        assert abs(float(blas.snrm2(n, dx, ix)) - wn) <= 2 * n * E * wn
        assert blas.isamax(n, dx, ix) == c["isamax"]
        assert sha(blas.sscal(n, a, dx, ix).to_host()) == c["sscal"]

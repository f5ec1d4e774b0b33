"""Golden fixtures, two sets with the same layout:

  "reference"  tests/golden/ref_golden_{f32,f64}.json -- outputs of the REFERENCE'S OWN SOURCE
               (src/Numerical/BLAS/{Types,Util,Single}.hs) evaluated by oracle/hs_eval.py in the build container
               (tests/golden/make_ref_golden.py): this is what pins the oracle and the CUDA path to the reference;
  "oracle"     tests/golden/level1_golden{,_f64}.json -- outputs of oracle/lblas_oracle.c (make_golden.py): guards
               the checker against drift.

CPU: the oracle must reproduce both bit for bit -- reductions included, it is as sequential as the reference.
GPU: the CUDA path must match them -- isamax and every elementwise routine bit for bit (NaN matches NaN),
the reductions within 2*n*eps*sum|x_i*y_i| -- including the non-finite / signed-zero / denormal cases
the reference's own generators never produce ("spiced" input sets) and every increment sign, 0 included.
"""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_NAMES = {("oracle", np.float32): "level1_golden.json", ("oracle", np.float64): "level1_golden_f64.json",
          ("reference", np.float32): "ref_golden_f32.json", ("reference", np.float64): "ref_golden_f64.json"}
FILES = {k: json.load(open(os.path.join(HERE, "golden", v))) for k, v in _NAMES.items()}
DTYPES = [np.float32, np.float64]
SETS = [pytest.param(w, d, id=f"{w}-{np.dtype(d).name}") for w in ("reference", "oracle") for d in DTYPES]
DT = np.float32          # set by each test: the instance (Float / Double) the file was made for
G = FILES[("oracle", np.float32)]
MGR_FIELDS = {"FLAGNEG2": (), "FLAGNEG1": ("d1", "d2", "x1", "h11", "h12", "h21", "h22"),
              "FLAG0": ("d1", "d2", "x1", "h12", "h21"), "FLAG1": ("d1", "d2", "x1", "h11", "h22")}      # Types.hs:15-20
MGR_FLAG = {"FLAGNEG2": -2, "FLAGNEG1": -1, "FLAG0": 0, "FLAG1": 1}


def use(dtype, which="oracle"):
    global DT, G, EPS
    DT, G = dtype, FILES[(which, dtype)]
    EPS = 2.0 ** -24 if dtype == np.float32 else 2.0 ** -53


def check_scalars(srotg, srotmg):
    """rotg / rotmg against the current file's scalar cases (the two sets store rotmg differently)"""
    for s in G["scalars"]:
        a = f32(s["args"])
        assert same(list(srotg(a[0], a[1])), f32(s["rotg"])), ("rotg", list(a))
        m = srotmg(*a)
        if isinstance(s["rotmg"][0], str):       # reference set: constructor name + the constructor's own fields
            con = s["rotmg"][0]
            assert m.flag == MGR_FLAG[con], ("rotmg", list(a))
            assert same([getattr(m, f) for f in MGR_FIELDS[con]], f32(s["rotmg"][1:])), ("rotmg", con, list(a))
        else:                                    # oracle set: flag + all seven fields (absent ones 0)
            assert m.flag == s["rotmg"][0]
            assert same([m.d1, m.d2, m.x1, m.h11, m.h12, m.h21, m.h22], f32(s["rotmg"][1:]))


EPS = 2.0 ** -24


def f32(u):
    """bit patterns -> values of the current instance's type"""
    return np.array(u, dtype=np.uint32 if DT == np.float32 else np.uint64).view(DT)


def same(got, want):
    got, want = np.asarray(got, dtype=DT), np.asarray(want, dtype=DT)
    U = np.uint32 if DT == np.float32 else np.uint64
    nan = np.isnan(want)
    return got.shape == want.shape and np.array_equal(np.isnan(got), nan) and \
        np.array_equal(got.view(U)[~nan], want.view(U)[~nan])


def sample(v, n, inc):
    first = 0 if inc > 0 else (1 - n) * inc
    return np.asarray(v, dtype=np.float64)[first + np.arange(n) * inc]


def run(impl, to_dev, to_host, case):
    n, ix, iy = case["n"], case["incx"], case["incy"]
    a, c, s = (f32([case[k]])[0] for k in ("a", "c", "s"))
    x, y = f32(case["x"]), f32(case["y"])
    dx, dy = to_dev(x), to_dev(y)
    outs = {
        "sdot": impl.sdot(n, dx, ix, dy, iy), "sdsdot": impl.sdsdot(n, a, dx, ix, dy, iy) if DT == np.float32 else None,
        "sasum": impl.sasum(n, dx, ix), "snrm2": impl.snrm2(n, dx, ix), "isamax": impl.isamax(n, dx, ix),
        "sscal": to_host(impl.sscal(n, a, dx, ix)), "scopy": to_host(impl.scopy(n, dx, ix, dy, iy)),
        "saxpy": to_host(impl.saxpy(n, a, dx, ix, dy, iy)),
        "sswap": [to_host(v) for v in impl.sswap(n, dx, ix, dy, iy)],
        "srot": [to_host(v) for v in impl.srot(n, dx, ix, dy, iy, c, s)],
    }
    return outs, x, y


@pytest.mark.parametrize("which", ["reference", "oracle"])
def test_readme_example_in_golden(oracle, which):
    use(np.float32, which)
    r = G["readme_example"]
    assert r["sdot"] == int(np.float32(32.0).view(np.uint32))       # README.md:57-59 of the reference
    assert oracle.sdot(r["n"], f32(r["x"]), 1, f32(r["y"]), 1).view(np.uint32) == r["sdot"]


@pytest.mark.parametrize("which,dtype", SETS)
def test_oracle_reproduces_golden(oracle, which, dtype):
    use(dtype, which)
    srotg, srotmg = (oracle.srotg, oracle.srotmg) if dtype == np.float32 else (oracle.drotg, oracle.drotmg)
    for case in G["cases"]:
        outs, x, y = run(oracle, lambda v: v, lambda v: v, case)
        want = case["outs"]
        for k in ("sdot", "sdsdot", "sasum", "snrm2"):
            if outs[k] is not None:
                assert same([outs[k]], f32(want[k])), (k, case["n"], case["incx"], case["incy"])
        assert outs["isamax"] == want["isamax"][0]
        for k in ("sscal", "scopy", "saxpy"):
            assert same(outs[k], f32(want[k])), k
        for k in ("sswap", "srot"):
            assert same(outs[k][0], f32(want[k][0])) and same(outs[k][1], f32(want[k][1])), k
        flags = oracle.ModGivensRot(int(f32(case["param"])[0]))      # srotm consumes only sparam
        p = f32(case["param"])
        flags = oracle.ModGivensRot(int(p[0]), h11=p[1], h21=p[2], h12=p[3], h22=p[4])
        mx, my = oracle.srotm(flags, case["n"], x, case["incx"], y, case["incy"])
        assert same(mx, f32(want["srotm"][0])) and same(my, f32(want["srotm"][1]))
    check_scalars(srotg, srotmg)


@pytest.mark.parametrize("which,dtype", SETS)
def test_product_host_scalars_match_golden(blas, which, dtype):
    use(dtype, which)
    srotg, srotmg = (blas.srotg, blas.srotmg) if dtype == np.float32 else (blas.drotg, blas.drotmg)
    check_scalars(srotg, srotmg)


@pytest.mark.gpu
@pytest.mark.parametrize("which,dtype", SETS)
def test_cuda_path_matches_golden(blas, ctx, which, dtype):
    import lambda_blas_b200 as lb
    use(dtype, which)
    big = 1e37 if dtype == np.float32 else 1e300
    for case in G["cases"]:
        n, ix, iy = case["n"], case["incx"], case["incy"]
        saxpy_like_inc0 = (ix == 0 or iy == 0)
        outs, x, y = run(blas, ctx.to_device, lambda v: v.to_host(), case)
        want = case["outs"]
        finite = np.all(np.isfinite(sample(x, n, ix))) and np.all(np.isfinite(sample(y, n, iy))) if n > 0 else True
        tag = (n, ix, iy, case["spiced"])
        assert outs["isamax"] == want["isamax"][0], tag
        for k in ("sscal", "scopy", "saxpy"):
            assert same(outs[k], f32(want[k])), (k,) + tag
        for k in ("sswap", "srot"):
            assert same(outs[k][0], f32(want[k][0])) and same(outs[k][1], f32(want[k][1])), (k,) + tag
        p = f32(case["param"])
        flags = lb.ModGivensRot(int(p[0]), h11=p[1], h21=p[2], h12=p[3], h22=p[4])
        mx, my = blas.srotm(flags, n, ctx.to_device(x), ix, ctx.to_device(y), iy)
        assert same(mx.to_host(), f32(want["srotm"][0])) and same(my.to_host(), f32(want["srotm"][1])), ("srotm",) + tag
        if not finite:
            continue        # reductions over Inf/NaN inputs: order-dependent (Inf - Inf), outside the reference's domain
        if n > 0:
            prod = np.sum(np.abs(sample(x, n, ix) * sample(y, n, iy)))
            with np.errstate(over="ignore"):
                risky = prod > big or np.sum(sample(x, n, ix) ** 2) > big
            if risky or not np.isfinite(prod):
                continue    # partial sums may overflow in one summation order and not in another
            assert abs(float(outs["sdot"]) - float(f32(want["sdot"])[0])) <= 2 * n * EPS * prod, ("sdot",) + tag
            if dtype == np.float32:
                w = float(f32(want["sdsdot"])[0])
                assert abs(float(outs["sdsdot"]) - w) <= float(np.spacing(np.float32(abs(w)))), ("sdsdot",) + tag
        else:
            assert same([outs["sdot"]], f32(want["sdot"]))
            if dtype == np.float32:
                assert same([outs["sdsdot"]], f32(want["sdsdot"]))
        ws, wn = float(f32(want["sasum"])[0]), float(f32(want["snrm2"])[0])
        if n > 0 and ix > 0:
            assert abs(float(outs["sasum"]) - ws) <= 2 * n * EPS * np.sum(np.abs(sample(x, n, ix))), ("sasum",) + tag
            assert abs(float(outs["snrm2"]) - wn) <= 2 * n * EPS * wn, ("snrm2",) + tag
        else:
            assert float(outs["sasum"]) == ws and float(outs["snrm2"]) == wn

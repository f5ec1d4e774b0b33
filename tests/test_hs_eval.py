"""oracle/hs_eval.py -- the evaluator that runs the reference's Haskell source -- is itself a checker, so it is
tested: the language subset on small modules with known answers, the restated `base` / `vector` primitives on
their corner cases, and (in the build container, where /root/reference exists) that evaluating the reference's
source as it lies there reproduces the committed fixtures and agrees with the C oracle on fresh random inputs.
"""
import json
import os

import numpy as np
import pytest

from oracle import hs_eval as H
from tests.gen import gen_nice_float, nice_scalar, span

HAVE_REF = os.path.isdir(H.REFERENCE_SRC)
F32, F64 = np.float32, np.float64


def module(body):
    it = H.Interp()
    it.load_source("module T where\n" + body)
    return it


def test_layout_guards_where_let_case():
    it = module('''
fact :: Int -> Int
fact !n
    | n < 1     = 1
    | otherwise = n * fact (n-1)

classify x = case compare x 0 of
    LT -> neg
    EQ -> 0
    GT -> let a = 1
              b = a + 1
          in  a + b
    where
    neg = negate 1

poly a b =
    let r = s * 2
        s = a + b
    in  if (r > 10)
          then r
          else helper r
    where
    helper v = v + k
        where
        k = 100
''')
    assert it.call("fact", 10, ty=int) == 3628800
    assert [it.call("classify", v, ty=int) for v in (-5, 0, 7)] == [-1, 0, 3]
    assert it.call("poly", 1, 2, ty=int) == 106 and it.call("poly", 4, 5, ty=int) == 18


def test_sections_lambdas_composition_tuples_records_views():
    it = module('''
data P a = NONE | TWO { px, py :: !a } | ONE { px :: !a }
twice f = f . f
apply3 = twice (*3) $ 2
sub = (\\ x y -> x - y) 10 3
flipsec v = (v `minus`) 1
    where minus p q = p - q
pair (a,b) = (b,a)
mk px py = TWO { .. }
mk2 = ONE { px = 5 }
sel p = case p of
    NONE -> 0
    TWO u w -> u + w
    ONE u -> u
viaview (negate -> m) z@(_,q) = (m, q, z)
neglit (-1) = 100
neglit _ = 0
''')
    assert it.call("apply3", ty=int) == 18 and it.call("sub", ty=int) == 7 and it.call("flipsec", 9, ty=int) == 8
    assert it.call("pair", (1, 2), ty=int) == (2, 1)
    assert it.call("sel", it.call("mk", 3, 4, ty=int), ty=int) == 7 and it.call("sel", it.call("mk2", ty=int), ty=int) == 5
    assert it.call("viaview", 5, (7, 8), ty=int) == (-5, 8, (7, 8))
    assert it.call("neglit", -1, ty=int) == 100 and it.call("neglit", 1, ty=int) == 0


def test_laziness_and_out_of_order_bindings():
    it = module('''
f x = let a = b + 1
          b = x * 2
          unused = boom 1
      in  a
boom n = boom (n+1)
g c = if c then 1 else boom 0
k = fst (3, boom 0)
''')
    assert it.call("f", 5, ty=int) == 11 and it.call("g", True, ty=int) == 1 and it.call("k", ty=int) == 3


def test_float_arithmetic_is_single_rounded_ieee():
    it = module('''
add x y = x + y
lit = 16777217
tenth = 0.1
sq x = x^(2::Int)
cube x = x^(3::Int)
half x = x / 2
widen (F# f) = case float2Double# f of
    d -> D# d
''')
    assert it.call("add", F32(0.1), F32(0.2), ty=F32) == F32(0.1) + F32(0.2)
    assert it.call("add", F32(16777216), H.Lit(1), ty=F32) == F32(16777216)            # the literal takes the operand's type
    assert it.call("lit", ty=F32) == F32(16777216) and it.call("lit", ty=F64) == 16777217.0
    assert it.call("tenth", ty=F32).view(np.uint32) == 0x3DCCCCCD                # fromRational, correctly rounded
    x = F32(1.1)
    assert it.call("sq", x, ty=F32) == x * x and it.call("cube", x, ty=F32) == (x * x) * x
    assert it.call("half", F64(3), ty=F64) == 1.5
    assert type(it.call("widen", F32(0.1))) is np.float64 and it.call("widen", F32(0.1)) == float(F32(0.1))
    with pytest.raises(H.HsError):
        it.call("add", F32(1), F64(1))                                          # no implicit conversions in Haskell


def test_literal_rounding_does_not_round_twice():
    # 1 + 2^-24 + 2^-60 rounds to 1 + 2^-23 in binary32, but to 1.0 if it is first rounded to binary64 (-> 1 + 2^-24, a tie)
    q = 1 + H.Fraction(1, 2 ** 24) + H.Fraction(1, 2 ** 60)
    assert H.to_type(H.Lit(q, True), F32) == np.nextafter(F32(1), F32(2))


def test_base_primitives_on_corner_cases():
    it = module('''
cmp x y = compare x y
a x = abs x
s x = signum x
''')
    nan = F32(np.nan)
    assert it.call("cmp", nan, F32(1)).name == "GT" and it.call("cmp", F32(1), nan).name == "GT"
    assert it.call("cmp", F32(-0.0), F32(0.0)).name == "EQ" and it.call("cmp", F32(1), F32(2)).name == "LT"
    assert np.signbit(it.call("a", F32(-0.0))) == False and it.call("a", F32(-3)) == 3       # noqa: E712
    assert np.isnan(it.call("a", nan))
    assert it.call("s", F32(-2)) == -1 and it.call("s", F32(5)) == 1 and np.signbit(it.call("s", F32(-0.0)))


def test_vector_combinators_follow_vector_0_11():
    it = module('''
first v = V.maxIndexBy (comparing abs) v
sumit v = V.foldl' (+) 0 v
upd v = V.unsafeUpdate_ v (unstream $ fromStream (Stream go 0) (Exact 3)) (V.replicate 3 9)
    where go !ix = return $ Yield ix ix
rev v = V.reverse $ V.unsafeTake 2 v
''')
    v = lambda a: it.vector(np.array(a, F32))      # noqa: E731
    assert it.call("first", v([1, -3, 3, 2])) == 1                  # the first maximum stays
    assert it.call("first", v([np.nan, 5, 1])) == 0                 # a NaN at the front is never displaced ...
    assert it.call("first", v([1, np.nan, 5])) == 2                 # ... and never displaces
    assert it.call("sumit", v([1e8, 1, -1e8])) == F32(0)            # strictly left to right: (1e8 + 1) - 1e8 = 0 in binary32
    assert list(it.call("upd", v([1, 2, 3]))) == [9, 2, 3]          # an endless index stream ends with the values
    assert list(it.call("rev", v([1, 2, 3]))) == [2, 1]
    with pytest.raises(H.HsError):
        it.call("rev", v([1]))                                      # unsafeTake beyond the vector: undefined in the reference


needs_ref = pytest.mark.skipif(not HAVE_REF, reason="/root/reference is only present in the build container")


@needs_ref
def test_reference_readme_example():
    R = H.Reference()
    assert R.dot(3, np.array([1, 2, 3], F32), 1, np.array([4, 5, 6], F32), 1) == 32.0         # README.md:57-59


@needs_ref
@pytest.mark.parametrize("dt,fname", [(F32, "ref_golden_f32.json"), (F64, "ref_golden_f64.json")])
def test_committed_fixtures_are_what_the_reference_source_evaluates_to(dt, fname):
    """every third input set and scalar case, re-evaluated from /root/reference as it lies there"""
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fname)))
    U = np.uint32 if dt == F32 else np.uint64
    val = lambda u: np.array(u, dtype=U).view(dt)      # noqa: E731
    bits = lambda a: [int(b) for b in np.atleast_1d(np.asarray(a, dtype=dt)).view(U)]      # noqa: E731
    R = H.Reference()
    for c in G["cases"][::3]:
        n, ix, iy, x, y = c["n"], c["incx"], c["incy"], val(c["x"]), val(c["y"])
        a, cc, s = (val([c[k]])[0] for k in ("a", "c", "s"))
        o = c["outs"]
        assert bits(R.dot(n, x, ix, y, iy)) == o["sdot"] and bits(R.asum(n, x, ix)) == o["sasum"]
        assert bits(R.nrm2(n, x, ix)) == o["snrm2"] and [R.iamax(n, x, ix)] == o["isamax"]
        assert bits(R.scal(n, a, x, ix)) == o["sscal"] and bits(R.copy(n, x, ix, y, iy)) == o["scopy"]
        assert bits(R.axpy(n, a, x, ix, y, iy)) == o["saxpy"]
        assert [bits(v) for v in R.swap(n, x, ix, y, iy)] == o["sswap"]
        assert [bits(v) for v in R.rot(n, x, ix, y, iy, cc, s)] == o["srot"]
        if dt == F32:
            assert bits(R.sdsdot(n, a, x, ix, y, iy)) == o["sdsdot"]
    for sc in G["scalars"][::3]:
        a = val(sc["args"])
        assert bits(list(R.rotg(a[0], a[1], ty=dt))) == sc["rotg"]
        con, f = R.rotmg(*a, ty=dt)
        assert [con] + bits([f[k] for k in H.MGR_FIELDS[con]]) == sc["rotmg"]


@needs_ref
@pytest.mark.parametrize("dt", [F32, F64])
def test_c_oracle_equals_reference_source_on_fresh_inputs(oracle, dt):
    """the differential the fixtures freeze, on other seeds: oracle/lblas_oracle.c == the reference's source, bit for bit"""
    R = H.Reference()
    rng = np.random.default_rng(777 if dt == F32 else 778)
    U = np.uint32 if dt == F32 else np.uint64

    def same(a, b):
        a, b = np.atleast_1d(np.asarray(a, dtype=dt)), np.atleast_1d(np.asarray(b, dtype=dt))
        nan = np.isnan(a)
        return a.shape == b.shape and np.array_equal(nan, np.isnan(b)) and np.array_equal(a.view(U)[~nan], b.view(U)[~nan])

    for it in range(150):
        n = int(rng.integers(0, 40))
        ix, iy = int(rng.integers(-5, 6)), int(rng.integers(-5, 6))
        x = gen_nice_float(rng, span(max(n, 1), ix, int(rng.integers(0, 3))), dt)
        y = gen_nice_float(rng, span(max(n, 1), iy, int(rng.integers(0, 3))), dt)
        if it % 5 == 0 and x.size > 1:
            x[int(rng.integers(0, x.size))] = rng.choice([np.nan, np.inf, -0.0, 1e-30, 1e30])
        a = nice_scalar(rng, dt)
        tag = (n, ix, iy)
        assert same(R.dot(n, x, ix, y, iy), oracle.dot(n, x, ix, y, iy)), tag
        assert same(R.asum(n, x, ix), oracle.asum(n, x, ix)) and same(R.nrm2(n, x, ix), oracle.nrm2(n, x, ix)), tag
        assert R.iamax(n, x, ix) == oracle.iamax(n, x, ix), tag
        assert same(R.scal(n, a, x, ix), oracle.scal(n, a, x, ix)) and same(R.copy(n, x, ix, y, iy), oracle.copy(n, x, ix, y, iy)), tag
        assert same(R.axpy(n, a, x, ix, y, iy), oracle.axpy(n, a, x, ix, y, iy)), tag
        for r, o in zip(R.swap(n, x, ix, y, iy), oracle.swap(n, x, ix, y, iy)):
            assert same(r, o), tag
        c, s = dt(np.cos(0.3)), dt(np.sin(0.3))
        for r, o in zip(R.rot(n, x, ix, y, iy, c, s), oracle.rot(n, x, ix, y, iy, c, s)):
            assert same(r, o), tag
        if dt == F32:
            assert same(R.sdsdot(n, a, x, ix, y, iy), oracle.sdsdot(n, a, x, ix, y, iy)), tag
        margs = list(rng.uniform(-4, 4, 4).astype(dt)) if it % 2 else [nice_scalar(rng, dt) for _ in range(4)]
        con, f = R.rotmg(*margs, ty=dt)
        m = (oracle.srotmg if dt == F32 else oracle.drotmg)(*margs)
        assert m.flag == H.MGR_FLAG[con] and same([f[k] for k in H.MGR_FIELDS[con]], [getattr(m, k) for k in H.MGR_FIELDS[con]]), margs
        for r, o in zip(R.rotm((con, f), n, x, ix, y, iy), oracle.rotm(m, n, x, ix, y, iy)):
            assert same(r, o), tag
        g = (oracle.srotg if dt == F32 else oracle.drotg)(margs[0], margs[1])
        assert same(list(R.rotg(margs[0], margs[1], ty=dt)), list(g)), margs


@needs_ref
@pytest.mark.parametrize("dt", [F32, F64])
def test_c_oracle_equals_reference_source_on_arbitrary_bit_patterns(oracle, dt):
    """elements drawn from ALL bit patterns (NaN payloads, infinities, denormals, signed zeros), n from 0, increments -3..3"""
    R = H.Reference()
    rng = np.random.default_rng(2026 if dt == F32 else 2027)
    U = np.uint32 if dt == F32 else np.uint64

    def same(a, b):
        a, b = np.atleast_1d(np.asarray(a, dtype=dt)), np.atleast_1d(np.asarray(b, dtype=dt))
        nan = np.isnan(a)
        return a.shape == b.shape and np.array_equal(nan, np.isnan(b)) and np.array_equal(a.view(U)[~nan], b.view(U)[~nan])

    def vec(n):
        kind = int(rng.integers(3))
        if kind == 0:
            return rng.integers(0, 2 ** 32 if dt == F32 else 2 ** 63, n, dtype=np.uint64).astype(U).view(dt).copy()
        if kind == 1:
            with np.errstate(over="ignore"):
                return (rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n)).astype(dt)
        return rng.choice(np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 1e-40, 3e38, 2.0 ** 60, 2.0 ** -70]), n).astype(dt)

    for _ in range(120):
        n = int(rng.integers(0, 14))
        ix, iy = int(rng.integers(-3, 4)), int(rng.integers(-3, 4))
        x, y = vec(span(max(n, 1), ix, int(rng.integers(0, 2)))), vec(span(max(n, 1), iy, int(rng.integers(0, 2))))
        a = vec(1)[0]
        c, s = vec(2)
        tag = (n, ix, iy)
        with np.errstate(all="ignore"):
            assert same(R.dot(n, x, ix, y, iy), oracle.dot(n, x, ix, y, iy)), tag
            assert same(R.asum(n, x, ix), oracle.asum(n, x, ix)) and same(R.nrm2(n, x, ix), oracle.nrm2(n, x, ix)), tag
            assert R.iamax(n, x, ix) == oracle.iamax(n, x, ix), tag
            assert same(R.scal(n, a, x, ix), oracle.scal(n, a, x, ix)) and same(R.copy(n, x, ix, y, iy), oracle.copy(n, x, ix, y, iy)), tag
            assert same(R.axpy(n, a, x, ix, y, iy), oracle.axpy(n, a, x, ix, y, iy)), tag
            for r, o in zip(R.swap(n, x, ix, y, iy), oracle.swap(n, x, ix, y, iy)):
                assert same(r, o), tag
            for r, o in zip(R.rot(n, x, ix, y, iy, c, s), oracle.rot(n, x, ix, y, iy, c, s)):
                assert same(r, o), tag
            if dt == F32:
                assert same(R.sdsdot(n, a, x, ix, y, iy), oracle.sdsdot(n, a, x, ix, y, iy)), tag

"""The C ABI library loads, exports every symbol include/lbk.h declares, and its host-only entry points
(srotg, srotmg, shard partition, record fold) behave -- all without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests.gen import nice_scalar

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "lbk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lbk_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(blas):
    from lambda_blas_b200 import _ffi
    L = _ffi.lib()
    names = declared_symbols()
    assert len(names) >= 85
    for n in names:
        assert hasattr(L, n), f"liblbk.so does not export {n}"
    assert sorted(_ffi.PROTOTYPES) == names, "the ctypes binding and the header disagree"
    assert L.lbk_version() == 200


def test_no_cpu_fallback(blas):
    """Without a CUDA device the product refuses to run (it never routes through the oracle)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(blas.LbkError):
        blas.Context(0)
    src = "".join(open(os.path.join(ROOT, "lambda_blas_b200", f)).read()
                  for f in os.listdir(os.path.join(ROOT, "lambda_blas_b200")) if f.endswith(".py"))
    assert "oracle" not in src.replace("oracle binding", "").replace("test oracle", ""), "the product must not import the oracle"


def bits(a):
    return np.asarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_host_scalars_bit_exact(blas, oracle, dtype):      # Main.hs:31-34 (rotg, rotmg at Float and Double): product code on the CPU
    rng = np.random.default_rng(5)
    seen = set()
    U = np.uint32 if dtype == np.float32 else np.uint64
    bits = lambda a: np.asarray(a, dtype=dtype).view(U)  # noqa: E731
    p = "s" if dtype == np.float32 else "d"
    rotg, rotmg, orotg, orotmg = (getattr(m, p + n) for m in (blas, oracle) for n in ("rotg", "rotmg"))
    for it in range(6000):
        # the reference's domain (genNiceFloat / genNiceDouble); at Double its magnitudes (~1e15) always trip
        # checkscale, so 1000 moderate cases are added to reach the FLAG0 / FLAG1 constructors there too
        a = [nice_scalar(rng, dtype) for _ in range(4)] if it < 5000 else list(rng.uniform(-4, 4, 4).astype(dtype))
        assert np.array_equal(bits(list(rotg(a[0], a[1]))), bits(list(orotg(a[0], a[1]))))
        g, w = rotmg(*a), orotmg(*a)
        seen.add(g.flag)
        assert g.flag == w.flag
        assert np.array_equal(bits([g.d1, g.d2, g.x1, g.h11, g.h12, g.h21, g.h22]), bits(list(w)[1:]))
        assert np.array_equal(bits(g.sparam(dtype)), bits(w.sparam(dtype)))
    assert seen == {-2, -1, 0, 1}
    assert blas.srotmg(1.5, 0.75, 2.0, 0.5).constructor == "FLAG0"
    assert list(blas.srotg(0, 0)) == [0, 0, 1, 0]          # Single.hs:281


def test_srotmg_sparam_output(blas):
    from lambda_blas_b200 import _ffi
    out, par = (C.c_float * 8)(), (C.c_float * 5)()
    assert _ffi.lib().lbk_srotmg(1.5, 0.75, 2.0, 0.5, out, par) == 0
    assert list(par) == [0.0, 0.0, out[6], out[5], 0.0]     # [flag,h11,h21,h12,h22], Foreign.hs:279
    assert _ffi.lib().lbk_srotmg(1.0, 0.0, 1.0, 1.0, out, par) == 0 and list(par) == [-2.0, 1.0, 0.0, 0.0, 1.0]


@pytest.mark.parametrize("n,w", [(0, 1), (1, 4), (7, 3), (10, 8), (1 << 28, 8), ((1 << 30) + 5, 7)])
def test_shard_partition_tiles_the_range(blas, n, w):
    from lambda_blas_b200 import sharded
    prev = 0
    for r in range(w):
        lo, hi = sharded.shard_range(n, w, r)
        assert lo == prev and hi >= lo and hi - lo in (n // w, n // w + 1)
        prev = hi
    assert prev == n


def test_record_fold(blas):
    """What the finish kernel computes from the ranks' records, stated on the host."""
    from lambda_blas_b200 import sharded
    from lambda_blas_b200._ffi import Record

    def rec(f=(0, 0, 0), key=0, idx=-1):
        return Record((C.c_double * 3)(*(list(f) + [0, 0, 0])[:3]), key, idx, 0)
    key = lambda v: int(np.float32(abs(v)).view(np.uint32))  # noqa: E731
    # sum in rank order
    assert sharded.combine_records(0, [rec((1.5,)), rec((2.25,)), rec((-0.75,))]) == 3.0
    # isamax: larger |x| wins, ties go to the lower global index, empty shards are ignored
    assert sharded.combine_records(2, [rec(key=key(3.0), idx=5), rec(key=key(-3.0), idx=1000), rec()]) == 5
    assert sharded.combine_records(2, [rec(key=key(1.0), idx=5), rec(key=key(-3.0), idx=1000)]) == 1000
    assert sharded.combine_records(2, [rec(), rec(key=0, idx=77)]) == 77
    assert sharded.combine_records(2, [rec(key=2 ** 64 - 1, idx=0), rec(key=key(np.inf), idx=9)]) == 0   # NaN at index 0 stays
    # nrm2: the three scaled sums add, then the Blue combine
    got = sharded.combine_records(1, [rec((0, 9.0, 0)), rec((0, 16.0, 0))])
    assert got == 5.0
    big = np.float32(2.0 ** 100)
    sb = np.float32(2.0 ** -76)
    got = sharded.combine_records(1, [rec((float((big * sb) ** 2), 0, 0)), rec((float((big * sb) ** 2), 0, 0))])
    assert abs(got / (float(big) * np.sqrt(2.0)) - 1) < 1e-6


def _header_arity():
    """name -> number of parameters, from the prototypes of include/lbk.h."""
    src = open(os.path.join(ROOT, "include", "lbk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for name, params in re.findall(r"\b(lbk_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        params = params.strip()
        out[name] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def test_haskell_shim_binds_declared_symbols_with_matching_arity():
    """hs/Numerical/BLAS/Device.hs cannot be compiled here (no GHC): at least every `foreign import ccall` in it must
    name a symbol of include/lbk.h, with as many arguments as the C prototype takes, and cover both instances."""
    hs = open(os.path.join(ROOT, "hs", "Numerical", "BLAS", "Device.hs")).read()
    arity = _header_arity()
    imports = re.findall(r'foreign import ccall\s+(?:unsafe\s+)?"(&?)(lbk_[a-z0-9_]+)"\s+\w+\s*::\s*(.*)', hs)
    assert len(imports) >= 50
    for addr, name, ty in imports:
        assert name in arity, f"Device.hs imports {name}, which include/lbk.h does not declare"
        if addr:
            continue                                  # an address import (the finaliser), not a call
        nargs = len(re.split(r"\s->\s", ty)) - 1      # every "->" outside parentheses separates an argument
        assert "(" not in ty.replace("Ptr (V Float)", "").replace("Ptr (V Double)", "").replace("Ptr (Ptr LbkCtx)", "").replace("Ptr (Ptr LbkGraph)", ""), ty
        assert nargs == arity[name], f"{name}: Device.hs passes {nargs} arguments, include/lbk.h takes {arity[name]}"
    names = {n for _, n, _ in imports}
    for r in ("dot", "asum", "nrm2", "axpy", "scal", "copy", "swap", "rot", "rotm", "rotg", "rotmg",
              "axpy_oop", "scal_oop", "copy_oop", "swap_oop", "rot_oop", "rotm_oop"):
        assert f"lbk_s{r}" in names and f"lbk_d{r}" in names, r
    assert {"lbk_isamax", "lbk_idamax", "lbk_sdsdot", "lbk_vec_alloc", "lbk_vec_alloc_f64", "lbk_vec_free"} <= names


def test_shard_span_partitions_the_logical_range(blas):
    """Positive increments on shards: the ranks' (first, count) tile [0, n) in rank order and every element's
    position i*inc falls inside the owning rank's block at the reported offset."""
    from lambda_blas_b200 import sharded
    rng = np.random.default_rng(3)
    for _ in range(300):
        world = int(rng.integers(1, 9)); inc = int(rng.integers(1, 9)); n = int(rng.integers(0, 200))
        glen = (1 + (n - 1) * inc if n else 0) + int(rng.integers(0, 20))
        nxt = 0
        for r in range(world):
            lo, hi = sharded.shard_range(glen, world, r)
            first, count, off = sharded.shard_span(glen, world, r, n, inc)
            if count:
                assert first == nxt and lo <= first * inc < hi and first * inc - lo == off
                assert lo <= (first + count - 1) * inc < hi
                nxt = first + count
            else:
                assert not any(lo <= i * inc < hi for i in range(n))
        assert nxt == n
    with pytest.raises(blas.LbkError):
        sharded.shard_span(10, 2, 0, 5, 3)          # 1+(n-1)*inc exceeds the vector


@pytest.mark.parametrize("src", ["example.cpp", "latency.cpp"])
def test_cpp_mirror_compiles_against_the_header(src, tmp_path):
    """cpp/lambda_blas.hpp (the compiled host layer) and its programs build and link against liblbk.so; running them
    needs a GPU and is tests/test_gpu_smoke_first.py::test_cpp_mirror_example."""
    import shutil
    import subprocess
    if not shutil.which("g++"):
        pytest.skip("no g++")
    lib = os.path.join(ROOT, "lambda_blas_b200")
    if not os.path.exists(os.path.join(lib, "liblbk.so")):
        pytest.skip("liblbk.so not built")
    exe = str(tmp_path / "prog")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", os.path.join(ROOT, "cpp", src), "-o", exe, "-L" + lib, "-llbk",
                        "-Wl,-rpath," + lib], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]

"""ctypes front end of the CPU oracle (oracle/lblas_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package lambda_blas_b200 never does.

The functions keep the reference's PURE call surface (Numerical.BLAS.Single,
src/Numerical/BLAS/Single.hs:84,113,129,147,165,201,222,421,441,484): vector results
are NEW numpy arrays of the destination's full length; inputs are never modified.

PARITY UNPINNED in the strict sense (no golden vectors in the reference, no GHC
here) -- see the header of lblas_oracle.c for what pins it instead.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import NamedTuple, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_I64 = C.c_int64
_FP = C.POINTER(C.c_float)


def build(force: bool = False) -> None:
    """Compile oracle/*.c with the committed Makefile (gcc, seconds)."""
    targets = ["liblblas_oracle.so", "libnetlib_l1.so"]
    stale = force
    for t, src in zip(targets, ["lblas_oracle.c", "netlib_l1.c"]):
        so, c = os.path.join(_HERE, t), os.path.join(_HERE, src)
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(c):
            stale = True
    if stale:
        subprocess.run(["make", "-C", _HERE, "-s", "-B", "all"], check=True)


def _load(name: str) -> C.CDLL:
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


_lib: Optional[C.CDLL] = None
_nl: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = _load("liblblas_oracle.so")
        L.lbo_sdot.restype = C.c_float
        L.lbo_sdot.argtypes = [_I64, _FP, _I64, _FP, _I64]
        L.lbo_sdsdot.restype = C.c_float
        L.lbo_sdsdot.argtypes = [_I64, C.c_float, _FP, _I64, _FP, _I64]
        for f in (L.lbo_sasum, L.lbo_snrm2):
            f.restype = C.c_float
            f.argtypes = [_I64, _FP, _I64]
        L.lbo_isamax.restype = _I64
        L.lbo_isamax.argtypes = [_I64, _FP, _I64]
        L.lbo_sscal.restype = None
        L.lbo_sscal.argtypes = [_I64, C.c_float, _FP, _I64, _I64, _FP]
        L.lbo_scopy.restype = None
        L.lbo_scopy.argtypes = [_I64, _FP, _I64, _FP, _I64, _I64, _FP]
        L.lbo_sswap.restype = None
        L.lbo_sswap.argtypes = [_I64, _FP, _I64, _I64, _FP, _I64, _I64, _FP, _FP]
        L.lbo_saxpy.restype = None
        L.lbo_saxpy.argtypes = [_I64, C.c_float, _FP, _I64, _FP, _I64, _I64, _FP]
        L.lbo_srot.restype = None
        L.lbo_srot.argtypes = [_I64, _FP, _I64, _I64, _FP, _I64, _I64, C.c_float, C.c_float, _FP, _FP]
        L.lbo_srotm.restype = None
        L.lbo_srotm.argtypes = [_FP, _I64, _FP, _I64, _I64, _FP, _I64, _I64, _FP, _FP]
        L.lbo_srotg.restype = None
        L.lbo_srotg.argtypes = [C.c_float, C.c_float, _FP]
        L.lbo_srotmg.restype = None
        L.lbo_srotmg.argtypes = [C.c_float] * 4 + [_FP]
        _lib = L
    return _lib


def netlib() -> C.CDLL:
    """The netlib 3.7.0 restatement (oracle/netlib_l1.c), in-place Fortran semantics."""
    global _nl
    if _nl is None:
        L = _load("libnetlib_l1.so")
        L.nl_sdot.restype = C.c_float
        L.nl_sdot.argtypes = [_I64, _FP, _I64, _FP, _I64]
        L.nl_sdsdot.restype = C.c_float
        L.nl_sdsdot.argtypes = [_I64, C.c_float, _FP, _I64, _FP, _I64]
        for f in (L.nl_sasum, L.nl_snrm2):
            f.restype = C.c_float
            f.argtypes = [_I64, _FP, _I64]
        L.nl_isamax.restype = _I64
        L.nl_isamax.argtypes = [_I64, _FP, _I64]
        L.nl_saxpy.restype = None
        L.nl_saxpy.argtypes = [_I64, C.c_float, _FP, _I64, _FP, _I64]
        L.nl_sscal.restype = None
        L.nl_sscal.argtypes = [_I64, C.c_float, _FP, _I64]
        for f in (L.nl_scopy, L.nl_sswap):
            f.restype = None
            f.argtypes = [_I64, _FP, _I64, _FP, _I64]
        L.nl_srot.restype = None
        L.nl_srot.argtypes = [_I64, _FP, _I64, _FP, _I64, C.c_float, C.c_float]
        L.nl_srotm.restype = None
        L.nl_srotm.argtypes = [_I64, _FP, _I64, _FP, _I64, _FP]
        L.nl_srotg.restype = None
        L.nl_srotg.argtypes = [_FP, _FP, _FP, _FP]
        L.nl_srotmg.restype = None
        L.nl_srotmg.argtypes = [_FP, _FP, _FP, C.c_float, _FP]
        _nl = L
    return _nl


def _f32(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 1:
        raise ValueError("vectors are one-dimensional")
    return a


def _p(a: np.ndarray):
    # a zero-length array still needs a valid (never dereferenced) pointer
    if a.size == 0:
        return C.cast(C.c_void_p(0), _FP)
    return a.ctypes.data_as(_FP)


# ---- ADTs of src/Numerical/BLAS/Types.hs:8,15-20 --------------------------------------

class GivensRot(NamedTuple):
    r: np.float32
    z: np.float32
    c: np.float32
    s: np.float32


class ModGivensRot(NamedTuple):
    """flag -2/-1/0/1 = FLAGNEG2 / FLAGNEG1 / FLAG0 / FLAG1; FLAGNEG2 carries no fields."""
    flag: int
    d1: np.float32 = np.float32(0)
    d2: np.float32 = np.float32(0)
    x1: np.float32 = np.float32(0)
    h11: np.float32 = np.float32(0)
    h12: np.float32 = np.float32(0)
    h21: np.float32 = np.float32(0)
    h22: np.float32 = np.float32(0)

    def sparam(self) -> np.ndarray:
        """BLAS sparam layout [flag,h11,h21,h12,h22], filled as test/Foreign.hs:276-280 does."""
        if self.flag == -2:
            return np.array([-2, 1, 0, 0, 1], dtype=np.float32)
        if self.flag == -1:
            return np.array([-1, self.h11, self.h21, self.h12, self.h22], dtype=np.float32)
        if self.flag == 0:
            return np.array([0, 0, self.h21, self.h12, 0], dtype=np.float32)
        return np.array([1, self.h11, 0, 0, self.h22], dtype=np.float32)


# ---- reductions -----------------------------------------------------------------------

def sdot(n: int, x, incx: int, y, incy: int) -> np.float32:
    x, y = _f32(x), _f32(y)
    return np.float32(lib().lbo_sdot(n, _p(x), incx, _p(y), incy))


def sdsdot(n: int, sb, x, incx: int, y, incy: int) -> np.float32:
    x, y = _f32(x), _f32(y)
    return np.float32(lib().lbo_sdsdot(n, float(np.float32(sb)), _p(x), incx, _p(y), incy))


def sasum(n: int, x, incx: int) -> np.float32:
    x = _f32(x)
    return np.float32(lib().lbo_sasum(n, _p(x), incx))


def snrm2(n: int, x, incx: int) -> np.float32:
    x = _f32(x)
    return np.float32(lib().lbo_snrm2(n, _p(x), incx))


def isamax(n: int, x, incx: int) -> int:
    x = _f32(x)
    return int(lib().lbo_isamax(n, _p(x), incx))


# ---- pure vector results --------------------------------------------------------------

def sscal(n: int, a, x, incx: int) -> np.ndarray:
    x = _f32(x)
    out = np.empty_like(x)
    lib().lbo_sscal(n, float(np.float32(a)), _p(x), x.size, incx, _p(out))
    return out


def scopy(n: int, x, incx: int, y, incy: int) -> np.ndarray:
    x, y = _f32(x), _f32(y)
    out = np.empty_like(y)
    lib().lbo_scopy(n, _p(x), incx, _p(y), y.size, incy, _p(out))
    return out


def sswap(n: int, x, incx: int, y, incy: int) -> Tuple[np.ndarray, np.ndarray]:
    x, y = _f32(x), _f32(y)
    xo, yo = np.empty_like(x), np.empty_like(y)
    lib().lbo_sswap(n, _p(x), x.size, incx, _p(y), y.size, incy, _p(xo), _p(yo))
    return xo, yo


def saxpy(n: int, a, x, incx: int, y, incy: int) -> np.ndarray:
    x, y = _f32(x), _f32(y)
    out = np.empty_like(y)
    lib().lbo_saxpy(n, float(np.float32(a)), _p(x), incx, _p(y), y.size, incy, _p(out))
    return out


def srot(n: int, x, incx: int, y, incy: int, c, s) -> Tuple[np.ndarray, np.ndarray]:
    x, y = _f32(x), _f32(y)
    xo, yo = np.empty_like(x), np.empty_like(y)
    lib().lbo_srot(n, _p(x), x.size, incx, _p(y), y.size, incy,
                   float(np.float32(c)), float(np.float32(s)), _p(xo), _p(yo))
    return xo, yo


def srotm(flags: ModGivensRot, n: int, x, incx: int, y, incy: int) -> Tuple[np.ndarray, np.ndarray]:
    x, y = _f32(x), _f32(y)
    xo, yo = np.empty_like(x), np.empty_like(y)
    p = flags.sparam()
    lib().lbo_srotm(_p(p), n, _p(x), x.size, incx, _p(y), y.size, incy, _p(xo), _p(yo))
    return xo, yo


# ---- host scalars -----------------------------------------------------------------------

def srotg(sa, sb) -> GivensRot:
    out = np.zeros(4, dtype=np.float32)
    lib().lbo_srotg(float(np.float32(sa)), float(np.float32(sb)), _p(out))
    return GivensRot(*out)


def srotmg(sd1, sd2, sx1, sy1) -> ModGivensRot:
    out = np.zeros(8, dtype=np.float32)
    lib().lbo_srotmg(*(float(np.float32(v)) for v in (sd1, sd2, sx1, sy1)), _p(out))
    flag = int(out[0])
    if flag == -2:
        return ModGivensRot(-2)
    return ModGivensRot(flag, *out[1:])


# ---- in-place forms, used ONLY by the CPU timing legs of bench.py -------------------------

def saxpy_timed_pure(n: int, a, x: np.ndarray, y: np.ndarray, out: np.ndarray) -> None:
    """The reference's own cost: clone all of y, then the scatter loop (Util.hs:134-142)."""
    lib().lbo_saxpy(n, float(np.float32(a)), _p(x), 1, _p(y), y.size, 1, _p(out))

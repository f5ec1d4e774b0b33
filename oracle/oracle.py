"""ctypes front end of the CPU oracle (oracle/lblas_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package lambda_blas_b200 never does.

The functions keep the reference's PURE call surface (Numerical.BLAS.Single,
src/Numerical/BLAS/Single.hs:84,113,129,147,165,201,222,421,441,484): vector results
are NEW numpy arrays of the destination's full length; inputs are never modified.

Parity is pinned to the reference's own SOURCE, evaluated by oracle/hs_eval.py (no GHC
here, so not to a compiled build): tests/golden/ref_golden_*.json -- see the header of
lblas_oracle.c.
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
_DP = C.POINTER(C.c_double)


def _bind(L: C.CDLL, pre: str, netlib_style: bool) -> None:
    """Prototypes of the Float (s/is) and Double (d/id) instances."""
    for p, ip, ct, pt in (("s", "is", C.c_float, _FP), ("d", "id", C.c_double, _DP)):
        g = lambda name: getattr(L, f"{pre}_{name}")  # noqa: E731
        g(f"{p}dot").restype = ct
        g(f"{p}dot").argtypes = [_I64, pt, _I64, pt, _I64]
        for n in (f"{p}asum", f"{p}nrm2"):
            g(n).restype = ct
            g(n).argtypes = [_I64, pt, _I64]
        g(f"{ip}amax").restype = _I64
        g(f"{ip}amax").argtypes = [_I64, pt, _I64]
        for n in ("axpy", "scal", "copy", "swap", "rot", "rotm", "rotg", "rotmg"):
            g(p + n).restype = None
        if netlib_style:          # in place, Fortran argument order
            g(f"{p}axpy").argtypes = [_I64, ct, pt, _I64, pt, _I64]
            g(f"{p}scal").argtypes = [_I64, ct, pt, _I64]
            g(f"{p}copy").argtypes = [_I64, pt, _I64, pt, _I64]
            g(f"{p}swap").argtypes = [_I64, pt, _I64, pt, _I64]
            g(f"{p}rot").argtypes = [_I64, pt, _I64, pt, _I64, ct, ct]
            g(f"{p}rotm").argtypes = [_I64, pt, _I64, pt, _I64, pt]
            g(f"{p}rotg").argtypes = [pt, pt, pt, pt]
            g(f"{p}rotmg").argtypes = [pt, pt, pt, ct, pt]
        else:                     # pure: explicit output buffers
            g(f"{p}scal").argtypes = [_I64, ct, pt, _I64, _I64, pt]
            g(f"{p}copy").argtypes = [_I64, pt, _I64, pt, _I64, _I64, pt]
            g(f"{p}swap").argtypes = [_I64, pt, _I64, _I64, pt, _I64, _I64, pt, pt]
            g(f"{p}axpy").argtypes = [_I64, ct, pt, _I64, pt, _I64, _I64, pt]
            g(f"{p}rot").argtypes = [_I64, pt, _I64, _I64, pt, _I64, _I64, ct, ct, pt, pt]
            g(f"{p}rotm").argtypes = [pt, _I64, pt, _I64, _I64, pt, _I64, _I64, pt, pt]
            g(f"{p}rotg").argtypes = [ct, ct, pt]
            g(f"{p}rotmg").argtypes = [ct] * 4 + [pt]
    sd = getattr(L, f"{pre}_sdsdot")
    sd.restype = C.c_float
    sd.argtypes = [_I64, C.c_float, _FP, _I64, _FP, _I64]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = _load("liblblas_oracle.so")
        _bind(L, "lbo", False)
        _lib = L
    return _lib


def netlib() -> C.CDLL:
    """The netlib 3.7.0 restatement (oracle/netlib_l1.c), in-place Fortran semantics."""
    global _nl
    if _nl is None:
        L = _load("libnetlib_l1.so")
        _bind(L, "nl", True)
        _nl = L
    return _nl


def _real(a) -> np.ndarray:
    """float64 stays float64 (the Double instance); everything else becomes float32 (the Float instance)."""
    dt = np.float64 if getattr(a, "dtype", None) == np.float64 else np.float32
    a = np.ascontiguousarray(a, dtype=dt)
    if a.ndim != 1:
        raise ValueError("vectors are one-dimensional")
    return a


def _f32(a) -> np.ndarray:
    return _real(np.asarray(a, dtype=np.float32))


def _p(a: np.ndarray):
    ptr = _DP if a.dtype == np.float64 else _FP
    # a zero-length array still needs a valid (never dereferenced) pointer
    if a.size == 0:
        return C.cast(C.c_void_p(0), ptr)
    return a.ctypes.data_as(ptr)


def _pair(x, y):
    x, y = _real(x), _real(y)
    if x.dtype != y.dtype:
        raise TypeError("both vectors must have the same precision")
    return x, y


def _pre(a: np.ndarray) -> str:
    return "d" if a.dtype == np.float64 else "s"


def _fn(a: np.ndarray, name: str):
    return getattr(lib(), f"lbo_{_pre(a)}{name}")


def _sc(a: np.ndarray, v) -> float:
    return float(a.dtype.type(v))


# ---- ADTs of src/Numerical/BLAS/Types.hs:8,15-20 --------------------------------------

class GivensRot(NamedTuple):
    r: float
    z: float
    c: float
    s: float


class ModGivensRot(NamedTuple):
    """flag -2/-1/0/1 = FLAGNEG2 / FLAGNEG1 / FLAG0 / FLAG1; FLAGNEG2 carries no fields."""
    flag: int
    d1: float = np.float32(0)
    d2: float = np.float32(0)
    x1: float = np.float32(0)
    h11: float = np.float32(0)
    h12: float = np.float32(0)
    h21: float = np.float32(0)
    h22: float = np.float32(0)

    def sparam(self, dtype=np.float32) -> np.ndarray:
        """BLAS sparam layout [flag,h11,h21,h12,h22], filled as test/Foreign.hs:276-280 does."""
        if self.flag == -2:
            return np.array([-2, 1, 0, 0, 1], dtype=dtype)
        if self.flag == -1:
            return np.array([-1, self.h11, self.h21, self.h12, self.h22], dtype=dtype)
        if self.flag == 0:
            return np.array([0, 0, self.h21, self.h12, 0], dtype=dtype)
        return np.array([1, self.h11, 0, 0, self.h22], dtype=dtype)


# ---- reductions (polymorphic: the precision of the arrays selects the instance) -----------

def dot(n: int, x, incx: int, y, incy: int):
    x, y = _pair(x, y)
    return x.dtype.type(_fn(x, "dot")(n, _p(x), incx, _p(y), incy))


def sdsdot(n: int, sb, x, incx: int, y, incy: int) -> np.float32:
    x, y = _f32(x), _f32(y)
    return np.float32(lib().lbo_sdsdot(n, float(np.float32(sb)), _p(x), incx, _p(y), incy))


def asum(n: int, x, incx: int):
    x = _real(x)
    return x.dtype.type(_fn(x, "asum")(n, _p(x), incx))


def nrm2(n: int, x, incx: int):
    x = _real(x)
    return x.dtype.type(_fn(x, "nrm2")(n, _p(x), incx))


def iamax(n: int, x, incx: int) -> int:
    x = _real(x)
    f = lib().lbo_idamax if x.dtype == np.float64 else lib().lbo_isamax
    return int(f(n, _p(x), incx))


# ---- pure vector results --------------------------------------------------------------

def scal(n: int, a, x, incx: int) -> np.ndarray:
    x = _real(x)
    out = np.empty_like(x)
    _fn(x, "scal")(n, _sc(x, a), _p(x), x.size, incx, _p(out))
    return out


def copy(n: int, x, incx: int, y, incy: int) -> np.ndarray:
    x, y = _pair(x, y)
    out = np.empty_like(y)
    _fn(x, "copy")(n, _p(x), incx, _p(y), y.size, incy, _p(out))
    return out


def swap(n: int, x, incx: int, y, incy: int) -> Tuple[np.ndarray, np.ndarray]:
    x, y = _pair(x, y)
    xo, yo = np.empty_like(x), np.empty_like(y)
    _fn(x, "swap")(n, _p(x), x.size, incx, _p(y), y.size, incy, _p(xo), _p(yo))
    return xo, yo


def axpy(n: int, a, x, incx: int, y, incy: int) -> np.ndarray:
    x, y = _pair(x, y)
    out = np.empty_like(y)
    _fn(x, "axpy")(n, _sc(x, a), _p(x), incx, _p(y), y.size, incy, _p(out))
    return out


def rot(n: int, x, incx: int, y, incy: int, c, s) -> Tuple[np.ndarray, np.ndarray]:
    x, y = _pair(x, y)
    xo, yo = np.empty_like(x), np.empty_like(y)
    _fn(x, "rot")(n, _p(x), x.size, incx, _p(y), y.size, incy, _sc(x, c), _sc(x, s), _p(xo), _p(yo))
    return xo, yo


def rotm(flags: ModGivensRot, n: int, x, incx: int, y, incy: int) -> Tuple[np.ndarray, np.ndarray]:
    x, y = _pair(x, y)
    xo, yo = np.empty_like(x), np.empty_like(y)
    p = flags.sparam(x.dtype)
    _fn(x, "rotm")(_p(p), n, _p(x), x.size, incx, _p(y), y.size, incy, _p(xo), _p(yo))
    return xo, yo


# ---- host scalars -----------------------------------------------------------------------

def _rotg(R, f, sa, sb) -> GivensRot:
    out = np.zeros(4, dtype=R)
    f(float(R(sa)), float(R(sb)), _p(out))
    return GivensRot(*out)


def _rotmg(R, f, sd1, sd2, sx1, sy1) -> ModGivensRot:
    out = np.zeros(8, dtype=R)
    f(*(float(R(v)) for v in (sd1, sd2, sx1, sy1)), _p(out))
    flag = int(out[0])
    if flag == -2:
        return ModGivensRot(-2)
    return ModGivensRot(flag, *out[1:])


def srotg(sa, sb) -> GivensRot:
    return _rotg(np.float32, lib().lbo_srotg, sa, sb)


def drotg(sa, sb) -> GivensRot:
    return _rotg(np.float64, lib().lbo_drotg, sa, sb)


def srotmg(sd1, sd2, sx1, sy1) -> ModGivensRot:
    return _rotmg(np.float32, lib().lbo_srotmg, sd1, sd2, sx1, sy1)


def drotmg(sd1, sd2, sx1, sy1) -> ModGivensRot:
    return _rotmg(np.float64, lib().lbo_drotmg, sd1, sd2, sx1, sy1)


# the reference's deprecated aliases (Single.hs:83-87, ...): both instances behind each polymorphic name
sdot = ddot = dot
sasum = dasum = asum
snrm2 = dnrm2 = nrm2
isamax = idamax = iamax
sscal = dscal = scal
scopy = dcopy = copy
sswap = dswap = swap
saxpy = daxpy = axpy
srot = drot = rot
srotm = drotm = rotm

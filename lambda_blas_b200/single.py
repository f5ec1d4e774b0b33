"""Numerical.BLAS.Single on the GPU: the reference's call surface over device vectors.

Same names, argument order and results as src/Numerical/BLAS/Single.hs:41-61 -- the polymorphic
names (`dot`, `axpy`, ...) and their deprecated aliases for Float (`sdot`, `saxpy`, ...) and Double
(`ddot`, `daxpy`, ...) -- with `Vector a` replaced by DVector.  The polymorphic names dispatch on the
vectors' element type exactly as the reference's type class does.  Like the reference the vector
routines are PURE: they return new vectors and leave their arguments alone (Util.hs:134-142); the
in-place forms the C ABI offers are in lambda_blas_b200.inplace.  Everything runs on the device
through liblbk.so; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import numpy as np

from . import _ffi
from .types import DVector, GivensRot, ModGivensRot

_L = _ffi.lib


def _ctx(*vs: DVector):
    c = vs[0].ctx
    for v in vs[1:]:
        if v.ctx is not c:
            raise ValueError("vectors live on different contexts")
    return c


def _p(v: DVector) -> str:
    """'s' or 'd': which instance of the routine the vector selects."""
    return "d" if v.dtype == np.float64 else "s"


def _r(v: DVector):
    return (np.float64, C.c_double) if v.dtype == np.float64 else (np.float32, C.c_float)


def _like(v: DVector) -> DVector:
    """An uninitialised vector with v's length, type and shard layout."""
    return v.like()


def _fn(v: DVector, name: str):
    return getattr(_L(), f"lbk_{_p(v)}{name}")


# ---- products / norms -----------------------------------------------------------------------------

def dot(n: int, sx: DVector, incx: int, sy: DVector, incy: int):
    """Single.hs:73-87."""
    (R, CT), c = _r(sx), _ctx(sx, sy)
    out = CT()
    _ffi.check(_fn(sx, "dot")(c.handle, n, sx.handle, incx, sy.handle, incy, C.byref(out)), c.handle)
    return R(out.value)


def sdsdot(n: int, sb, sx: DVector, incx: int, sy: DVector, incy: int) -> np.float32:
    """Single.hs:231-262 (Float only, like the reference)."""
    c, out = _ctx(sx, sy), C.c_float()
    _ffi.check(_L().lbk_sdsdot(c.handle, n, float(np.float32(sb)), sx.handle, incx, sy.handle, incy, C.byref(out)), c.handle)
    return np.float32(out.value)


def asum(n: int, sx: DVector, incx: int):
    """Single.hs:155-168."""
    (R, CT), c = _r(sx), sx.ctx
    out = CT()
    _ffi.check(_fn(sx, "asum")(c.handle, n, sx.handle, incx, C.byref(out)), c.handle)
    return R(out.value)


def nrm2(n: int, sx: DVector, incx: int):
    """Single.hs:174-204."""
    (R, CT), c = _r(sx), sx.ctx
    out = CT()
    _ffi.check(_fn(sx, "nrm2")(c.handle, n, sx.handle, incx, C.byref(out)), c.handle)
    return R(out.value)


def iamax(n: int, sx: DVector, incx: int) -> int:
    """Single.hs:211-225: 0-based logical index, -1 for n<1 or incx<1."""
    c, out = sx.ctx, C.c_int64()
    fn = _L().lbk_idamax if sx.dtype == np.float64 else _L().lbk_isamax
    _ffi.check(fn(c.handle, n, sx.handle, incx, C.byref(out)), c.handle)
    return int(out.value)


# ---- vector results (pure) ----------------------------------------------------------------------------

def scal(n: int, a, sx: DVector, incx: int) -> DVector:
    """Single.hs:101-116."""
    (R, _), c, out = _r(sx), sx.ctx, _like(sx)
    _ffi.check(_fn(sx, "scal_oop")(c.handle, n, float(R(a)), sx.handle, incx, out.handle), c.handle)
    return out


def copy(n: int, sx: DVector, incx: int, sy: DVector, incy: int) -> DVector:
    """Single.hs:119-132: sy with the sampled elements of sx written over it."""
    c, out = _ctx(sx, sy), _like(sy)
    _ffi.check(_fn(sx, "copy_oop")(c.handle, n, sx.handle, incx, sy.handle, incy, out.handle), c.handle)
    return out


def swap(n: int, sx: DVector, incx: int, sy: DVector, incy: int) -> Tuple[DVector, DVector]:
    """Single.hs:135-150."""
    c, xo, yo = _ctx(sx, sy), _like(sx), _like(sy)
    _ffi.check(_fn(sx, "swap_oop")(c.handle, n, sx.handle, incx, sy.handle, incy, xo.handle, yo.handle), c.handle)
    return xo, yo


def axpy(n: int, a, sx: DVector, incx: int, sy: DVector, incy: int) -> DVector:
    """Single.hs:430-444: a*x + y as a new vector of y's length."""
    (R, _), c, out = _r(sx), _ctx(sx, sy), _like(sy)
    _ffi.check(_fn(sx, "axpy_oop")(c.handle, n, float(R(a)), sx.handle, incx, sy.handle, incy, out.handle), c.handle)
    return out


def rot(n: int, u: DVector, incx: int, v: DVector, incy: int, c_, s_) -> Tuple[DVector, DVector]:
    """Single.hs:408-424."""
    (R, _), c, xo, yo = _r(u), _ctx(u, v), _like(u), _like(v)
    _ffi.check(_fn(u, "rot_oop")(c.handle, n, u.handle, incx, v.handle, incy, float(R(c_)), float(R(s_)),
                                 xo.handle, yo.handle), c.handle)
    return xo, yo


def rotm(flags: ModGivensRot, n: int, sx: DVector, incx: int, sy: DVector, incy: int) -> Tuple[DVector, DVector]:
    """Single.hs:463-487.  FLAGNEG2 hands back the arguments themselves, as the reference does (:473)."""
    if flags.flag == -2:
        return sx, sy
    (R, CT), c, xo, yo = _r(sx), _ctx(sx, sy), _like(sx), _like(sy)
    p = flags.sparam(R)
    _ffi.check(_fn(sx, "rotm_oop")(c.handle, n, sx.handle, incx, sy.handle, incy, p.ctypes.data_as(C.POINTER(CT)),
                                   xo.handle, yo.handle), c.handle)
    return xo, yo


# ---- host scalars ------------------------------------------------------------------------------------

def _rotg(R, CT, fn, sa, sb) -> GivensRot:
    out = (CT * 4)()
    _ffi.check(fn(float(R(sa)), float(R(sb)), out))
    return GivensRot(*(R(v) for v in out))


def _rotmg(R, CT, fn, sd1, sd2, sx1, sy1) -> ModGivensRot:
    out = (CT * 8)()
    _ffi.check(fn(*(float(R(v)) for v in (sd1, sd2, sx1, sy1)), out, None))
    flag = int(out[0])
    if flag == -2:
        return ModGivensRot(-2)
    return ModGivensRot(flag, *(R(v) for v in list(out)[1:]))


def srotg(sa, sb) -> GivensRot:
    """Single.hs:274-300 at Float."""
    return _rotg(np.float32, C.c_float, _L().lbk_srotg, sa, sb)


def drotg(sa, sb) -> GivensRot:
    """Single.hs:274-300 at Double."""
    return _rotg(np.float64, C.c_double, _L().lbk_drotg, sa, sb)


def srotmg(sd1, sd2, sx1, sy1) -> ModGivensRot:
    """Single.hs:318-357 at Float."""
    return _rotmg(np.float32, C.c_float, _L().lbk_srotmg, sd1, sd2, sx1, sy1)


def drotmg(sd1, sd2, sx1, sy1) -> ModGivensRot:
    """Single.hs:318-357 at Double."""
    return _rotmg(np.float64, C.c_double, _L().lbk_drotmg, sd1, sd2, sx1, sy1)


def rotg(sa, sb) -> GivensRot:
    """Polymorphic like the reference: float64 arguments select the Double instance."""
    return drotg(sa, sb) if isinstance(sa, np.float64) else srotg(sa, sb)


def rotmg(sd1, sd2, sx1, sy1) -> ModGivensRot:
    return drotmg(sd1, sd2, sx1, sy1) if isinstance(sd1, np.float64) else srotmg(sd1, sd2, sx1, sy1)


# the deprecated aliases of the reference (Single.hs:83-87, 112-116, ...): Float and Double instances
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

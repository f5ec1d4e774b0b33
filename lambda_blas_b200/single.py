"""Numerical.BLAS.Single on the GPU: the reference's call surface over device vectors.

Same names, argument order and results as src/Numerical/BLAS/Single.hs:41-61 -- the polymorphic
names (`dot`, `axpy`, ...) and their deprecated single-precision aliases (`sdot`, `saxpy`, ...) --
with `Vector Float` replaced by DVector.  Like the reference the vector routines are PURE: they
return new vectors and leave their arguments alone (Util.hs:134-142); the in-place forms the C
ABI offers are in lambda_blas_b200.inplace.  Everything runs on the device through liblbk.so;
there is no CPU path.
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


def _like(v: DVector) -> DVector:
    """An uninitialised vector with v's length and shard layout."""
    c = v.ctx
    return c.alloc_sharded(v.global_len) if v.global_len != len(v) or v.shard_offset else c.alloc(len(v))


# ---- products / norms -----------------------------------------------------------------------------

def dot(n: int, sx: DVector, incx: int, sy: DVector, incy: int) -> np.float32:
    """Single.hs:73-87."""
    c, out = _ctx(sx, sy), C.c_float()
    _ffi.check(_L().lbk_sdot(c.handle, n, sx.handle, incx, sy.handle, incy, C.byref(out)), c.handle)
    return np.float32(out.value)


def sdsdot(n: int, sb, sx: DVector, incx: int, sy: DVector, incy: int) -> np.float32:
    """Single.hs:231-262."""
    c, out = _ctx(sx, sy), C.c_float()
    _ffi.check(_L().lbk_sdsdot(c.handle, n, float(np.float32(sb)), sx.handle, incx, sy.handle, incy, C.byref(out)), c.handle)
    return np.float32(out.value)


def asum(n: int, sx: DVector, incx: int) -> np.float32:
    """Single.hs:155-168."""
    c, out = sx.ctx, C.c_float()
    _ffi.check(_L().lbk_sasum(c.handle, n, sx.handle, incx, C.byref(out)), c.handle)
    return np.float32(out.value)


def nrm2(n: int, sx: DVector, incx: int) -> np.float32:
    """Single.hs:174-204."""
    c, out = sx.ctx, C.c_float()
    _ffi.check(_L().lbk_snrm2(c.handle, n, sx.handle, incx, C.byref(out)), c.handle)
    return np.float32(out.value)


def iamax(n: int, sx: DVector, incx: int) -> int:
    """Single.hs:211-225: 0-based logical index, -1 for n<1 or incx<1."""
    c, out = sx.ctx, C.c_int64()
    _ffi.check(_L().lbk_isamax(c.handle, n, sx.handle, incx, C.byref(out)), c.handle)
    return int(out.value)


# ---- vector results (pure) ----------------------------------------------------------------------------

def scal(n: int, a, sx: DVector, incx: int) -> DVector:
    """Single.hs:101-116."""
    c, out = sx.ctx, _like(sx)
    _ffi.check(_L().lbk_sscal_oop(c.handle, n, float(np.float32(a)), sx.handle, incx, out.handle), c.handle)
    return out


def copy(n: int, sx: DVector, incx: int, sy: DVector, incy: int) -> DVector:
    """Single.hs:119-132: sy with the sampled elements of sx written over it."""
    c, out = _ctx(sx, sy), _like(sy)
    _ffi.check(_L().lbk_scopy_oop(c.handle, n, sx.handle, incx, sy.handle, incy, out.handle), c.handle)
    return out


def swap(n: int, sx: DVector, incx: int, sy: DVector, incy: int) -> Tuple[DVector, DVector]:
    """Single.hs:135-150."""
    c, xo, yo = _ctx(sx, sy), _like(sx), _like(sy)
    _ffi.check(_L().lbk_sswap_oop(c.handle, n, sx.handle, incx, sy.handle, incy, xo.handle, yo.handle), c.handle)
    return xo, yo


def axpy(n: int, a, sx: DVector, incx: int, sy: DVector, incy: int) -> DVector:
    """Single.hs:430-444: a*x + y as a new vector of y's length."""
    c, out = _ctx(sx, sy), _like(sy)
    _ffi.check(_L().lbk_saxpy_oop(c.handle, n, float(np.float32(a)), sx.handle, incx, sy.handle, incy, out.handle), c.handle)
    return out


def rot(n: int, u: DVector, incx: int, v: DVector, incy: int, c_, s_) -> Tuple[DVector, DVector]:
    """Single.hs:408-424."""
    c, xo, yo = _ctx(u, v), _like(u), _like(v)
    _ffi.check(_L().lbk_srot_oop(c.handle, n, u.handle, incx, v.handle, incy, float(np.float32(c_)), float(np.float32(s_)),
                                 xo.handle, yo.handle), c.handle)
    return xo, yo


def rotm(flags: ModGivensRot, n: int, sx: DVector, incx: int, sy: DVector, incy: int) -> Tuple[DVector, DVector]:
    """Single.hs:463-487.  FLAGNEG2 hands back the arguments themselves, as the reference does (:473)."""
    if flags.flag == -2:
        return sx, sy
    c, xo, yo = _ctx(sx, sy), _like(sx), _like(sy)
    p = flags.sparam()
    _ffi.check(_L().lbk_srotm_oop(c.handle, n, sx.handle, incx, sy.handle, incy, p.ctypes.data_as(C.POINTER(C.c_float)),
                                  xo.handle, yo.handle), c.handle)
    return xo, yo


# ---- host scalars ------------------------------------------------------------------------------------

def rotg(sa, sb) -> GivensRot:
    """Single.hs:274-300."""
    out = (C.c_float * 4)()
    _ffi.check(_L().lbk_srotg(float(np.float32(sa)), float(np.float32(sb)), out))
    return GivensRot(*(np.float32(v) for v in out))


def rotmg(sd1, sd2, sx1, sy1) -> ModGivensRot:
    """Single.hs:318-357."""
    out = (C.c_float * 8)()
    _ffi.check(_L().lbk_srotmg(*(float(np.float32(v)) for v in (sd1, sd2, sx1, sy1)), out, None))
    flag = int(out[0])
    if flag == -2:
        return ModGivensRot(-2)
    return ModGivensRot(flag, *(np.float32(v) for v in list(out)[1:]))


# the deprecated single-precision aliases of the reference (Single.hs:83-87, 112-114, ...)
sdot, sasum, snrm2, isamax = dot, asum, nrm2, iamax
sscal, scopy, sswap, saxpy, srot, srotm, srotg, srotmg = scal, copy, swap, axpy, rot, rotm, rotg, rotmg

"""In-place, enqueue-only forms of the elementwise routines (the raw C ABI, include/lbk.h).

The reference has no in-place forms (its vectors are immutable); these are what a caller that owns
its device buffers uses to avoid the clone, and what bench.py times: the algorithmic bytes per
element of SURVEY.md 8(d) (saxpy 12, sscal/scopy 8, sswap/srot/srotm 16) are for these.  Each
function serves both instances (Float / Double) and picks the s* or d* entry point from the vectors.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .types import DVector, ModGivensRot

_L = _ffi.lib


def _fn(v: DVector, name: str):
    return getattr(_L(), f"lbk_{'d' if v.dtype == np.float64 else 's'}{name}")


def _R(v: DVector):
    return np.float64 if v.dtype == np.float64 else np.float32


def saxpy_(n: int, a, x: DVector, incx: int, y: DVector, incy: int) -> None:
    _ffi.check(_fn(x, "axpy")(x.ctx.handle, n, float(_R(x)(a)), x.handle, incx, y.handle, incy), x.ctx.handle)


def sscal_(n: int, a, x: DVector, incx: int) -> None:
    _ffi.check(_fn(x, "scal")(x.ctx.handle, n, float(_R(x)(a)), x.handle, incx), x.ctx.handle)


def scopy_(n: int, x: DVector, incx: int, y: DVector, incy: int) -> None:
    _ffi.check(_fn(x, "copy")(x.ctx.handle, n, x.handle, incx, y.handle, incy), x.ctx.handle)


def sswap_(n: int, x: DVector, incx: int, y: DVector, incy: int) -> None:
    _ffi.check(_fn(x, "swap")(x.ctx.handle, n, x.handle, incx, y.handle, incy), x.ctx.handle)


def srot_(n: int, x: DVector, incx: int, y: DVector, incy: int, c, s) -> None:
    R = _R(x)
    _ffi.check(_fn(x, "rot")(x.ctx.handle, n, x.handle, incx, y.handle, incy, float(R(c)), float(R(s))), x.ctx.handle)


def srotm_(flags: ModGivensRot, n: int, x: DVector, incx: int, y: DVector, incy: int) -> None:
    R = _R(x)
    p = flags.sparam(R)
    ct = C.c_double if R is np.float64 else C.c_float
    _ffi.check(_fn(x, "rotm")(x.ctx.handle, n, x.handle, incx, y.handle, incy, p.ctypes.data_as(C.POINTER(ct))), x.ctx.handle)


daxpy_, dscal_, dcopy_, dswap_, drot_, drotm_ = saxpy_, sscal_, scopy_, sswap_, srot_, srotm_


# enqueue-only reductions: the result lands in out[0] (i?amax: the int64 in the first 8 bytes) on the stream
def sdot_async(n: int, x: DVector, incx: int, y: DVector, incy: int, out: DVector) -> None:
    _ffi.check(_fn(x, "dot_async")(x.ctx.handle, n, x.handle, incx, y.handle, incy, out.handle), x.ctx.handle)


def sdsdot_async(n: int, sb, x: DVector, incx: int, y: DVector, incy: int, out: DVector) -> None:
    _ffi.check(_L().lbk_sdsdot_async(x.ctx.handle, n, float(np.float32(sb)), x.handle, incx, y.handle, incy, out.handle), x.ctx.handle)


def sasum_async(n: int, x: DVector, incx: int, out: DVector) -> None:
    _ffi.check(_fn(x, "asum_async")(x.ctx.handle, n, x.handle, incx, out.handle), x.ctx.handle)


def snrm2_async(n: int, x: DVector, incx: int, out: DVector) -> None:
    _ffi.check(_fn(x, "nrm2_async")(x.ctx.handle, n, x.handle, incx, out.handle), x.ctx.handle)


def isamax_async(n: int, x: DVector, incx: int, out: DVector) -> None:
    fn = _L().lbk_idamax_async if x.dtype == np.float64 else _L().lbk_isamax_async
    _ffi.check(fn(x.ctx.handle, n, x.handle, incx, out.handle), x.ctx.handle)


ddot_async, dasum_async, dnrm2_async, idamax_async = sdot_async, sasum_async, snrm2_async, isamax_async


def saxpy_into(n: int, a, x: DVector, incx: int, y: DVector, incy: int, yout: DVector) -> None:
    """yout = the reference's pure axpy result (y untouched), enqueue only: lbk_?axpy_oop."""
    _ffi.check(_fn(x, "axpy_oop")(x.ctx.handle, n, float(_R(x)(a)), x.handle, incx, y.handle, incy, yout.handle), x.ctx.handle)

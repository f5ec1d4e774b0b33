"""In-place, enqueue-only forms of the elementwise routines (the raw C ABI, include/lbk.h).

The reference has no in-place forms (its vectors are immutable); these are what a caller that owns
its device buffers uses to avoid the clone, and what bench.py times: the algorithmic bytes per
element of SURVEY.md 8(d) (saxpy 12, sscal/scopy 8, sswap/srot/srotm 16) are for these.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi
from .types import DVector, ModGivensRot

_L = _ffi.lib


def saxpy_(n: int, a, x: DVector, incx: int, y: DVector, incy: int) -> None:
    _ffi.check(_L().lbk_saxpy(x.ctx.handle, n, float(np.float32(a)), x.handle, incx, y.handle, incy), x.ctx.handle)


def sscal_(n: int, a, x: DVector, incx: int) -> None:
    _ffi.check(_L().lbk_sscal(x.ctx.handle, n, float(np.float32(a)), x.handle, incx), x.ctx.handle)


def scopy_(n: int, x: DVector, incx: int, y: DVector, incy: int) -> None:
    _ffi.check(_L().lbk_scopy(x.ctx.handle, n, x.handle, incx, y.handle, incy), x.ctx.handle)


def sswap_(n: int, x: DVector, incx: int, y: DVector, incy: int) -> None:
    _ffi.check(_L().lbk_sswap(x.ctx.handle, n, x.handle, incx, y.handle, incy), x.ctx.handle)


def srot_(n: int, x: DVector, incx: int, y: DVector, incy: int, c, s) -> None:
    _ffi.check(_L().lbk_srot(x.ctx.handle, n, x.handle, incx, y.handle, incy, float(np.float32(c)), float(np.float32(s))), x.ctx.handle)


def srotm_(flags: ModGivensRot, n: int, x: DVector, incx: int, y: DVector, incy: int) -> None:
    p = flags.sparam()
    _ffi.check(_L().lbk_srotm(x.ctx.handle, n, x.handle, incx, y.handle, incy, p.ctypes.data_as(C.POINTER(C.c_float))), x.ctx.handle)


# enqueue-only reductions: the result lands in out[0] (isamax: the int64 in out[0..1]) on the stream
def sdot_async(n: int, x: DVector, incx: int, y: DVector, incy: int, out: DVector) -> None:
    _ffi.check(_L().lbk_sdot_async(x.ctx.handle, n, x.handle, incx, y.handle, incy, out.handle), x.ctx.handle)


def sasum_async(n: int, x: DVector, incx: int, out: DVector) -> None:
    _ffi.check(_L().lbk_sasum_async(x.ctx.handle, n, x.handle, incx, out.handle), x.ctx.handle)


def snrm2_async(n: int, x: DVector, incx: int, out: DVector) -> None:
    _ffi.check(_L().lbk_snrm2_async(x.ctx.handle, n, x.handle, incx, out.handle), x.ctx.handle)


def isamax_async(n: int, x: DVector, incx: int, out: DVector) -> None:
    _ffi.check(_L().lbk_isamax_async(x.ctx.handle, n, x.handle, incx, out.handle), x.ctx.handle)


def saxpy_into(n: int, a, x: DVector, incx: int, y: DVector, incy: int, yout: DVector) -> None:
    """yout = the reference's pure saxpy result (y untouched), enqueue only: lbk_saxpy_oop."""
    _ffi.check(_L().lbk_saxpy_oop(x.ctx.handle, n, float(np.float32(a)), x.handle, incx, y.handle, incy, yout.handle), x.ctx.handle)

"""The reference's own source (evaluated by oracle/hs_eval.py) behind the call surface of oracle/oracle.py, so that
a test written against the C restatement can run, unchanged, against the reference itself.  Build container only
(/root/reference must exist)."""
import numpy as np

from oracle import hs_eval
from oracle import oracle as _o


class ReferenceAsOracle:
    GivensRot, ModGivensRot = _o.GivensRot, _o.ModGivensRot

    def __init__(self):
        self.R = hs_eval.Reference()
        for name in ("dot", "sdsdot", "asum", "nrm2", "iamax", "scal", "copy", "swap", "axpy", "rot"):
            setattr(self, name, getattr(self.R, name))
        self.sdot = self.ddot = self.R.dot
        self.sasum = self.dasum = self.R.asum
        self.snrm2 = self.dnrm2 = self.R.nrm2
        self.isamax = self.idamax = self.R.iamax
        self.sscal = self.dscal = self.R.scal
        self.scopy = self.dcopy = self.R.copy
        self.sswap = self.dswap = self.R.swap
        self.saxpy = self.daxpy = self.R.axpy
        self.srot = self.drot = self.R.rot
        self.srotm = self.drotm = self.rotm

    @staticmethod
    def build():
        _o.build()

    @staticmethod
    def netlib():
        return _o.netlib()

    @staticmethod
    def _arr(a):
        a = np.asarray(a)
        return a if a.dtype in (np.float32, np.float64) else a.astype(np.float32)

    def _mgr(self, con, f, ty):
        z = ty(0)
        return _o.ModGivensRot(hs_eval.MGR_FLAG[con], *[f.get(k, z) for k in ("d1", "d2", "x1", "h11", "h12", "h21", "h22")]) \
            if con != "FLAGNEG2" else _o.ModGivensRot(-2)

    def rotm(self, flags, n, x, incx, y, incy):
        con = {v: k for k, v in hs_eval.MGR_FLAG.items()}[flags.flag]
        return self.R.rotm((con, {k: getattr(flags, k) for k in hs_eval.MGR_FIELDS[con]}), n, x, incx, y, incy)

    def srotg(self, a, b):
        return _o.GivensRot(*self.R.rotg(a, b, ty=np.float32))

    def drotg(self, a, b):
        return _o.GivensRot(*self.R.rotg(a, b, ty=np.float64))

    def srotmg(self, *args):
        return self._mgr(*self.R.rotmg(*args, ty=np.float32), np.float32)

    def drotmg(self, *args):
        return self._mgr(*self.R.rotmg(*args, ty=np.float64), np.float64)

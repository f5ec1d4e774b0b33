"""One launch of every kernel shape the headline capture (tools/profile_one.py) does not cover, for ncu:
general-stride map / reduce at the BASELINE.json configs[3] increments, the reversed-lane (REV) shapes, every Double
instance, srotm's FLAGNEG1 / FLAG1 functors, the differently-misaligned (one element per access) shapes -- and, for the
DRAM read/write-turnaround question, a pure-read, a pure-write and the mixed streams side by side.

    python tools/ncu_more.py [log2n] > gpurun_out/more_manifest.csv        # prints label,kernel_family,algorithmic_bytes,sector_bytes
    ncu --metrics <list> -k regex:'unit_kernel|strided_kernel' --csv --log-file gpurun_out/more.csv python tools/ncu_more.py

Launch order = manifest order (fills and copies that set operands up are other kernels and are filtered out by -k).
tools/ncu_join.py joins the two by order.
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402
from lambda_blas_b200 import inplace as ip  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
n = 1 << lg
GROUPS = os.environ.get("LBK_GROUPS", "strided,rotm,misaligned,streams,f64").split(",")
ctx = lb.Context(0)
ctx.set_pdl(False)                       # one kernel at a time: ncu serialises launches anyway
span = 1 + (n - 1) * 7 if "strided" in GROUPS else n
X, Y, out = ctx.alloc(span + 16), ctx.alloc(span + 16), ctx.alloc(4)
c, s = math.cos(0.3), math.sin(0.3)
rows = []


def sector_bytes(n_, inc, esz=4):
    """DRAM bytes one pass over n_ elements at increment inc touches, in 32-byte sectors."""
    a = abs(inc)
    if a == 0:
        return 32
    if esz * a <= 32:              # consecutive elements are at most a sector apart: every sector of the span is touched
        return 32 * ((esz * (1 + (n_ - 1) * a) + 31) // 32)
    return 32 * n_                 # each element in a sector of its own


def fresh():
    # same-kernel fills are fill_hash_kernel (not matched by -k)
    X.fill_hash(1); Y.fill_hash(2)


def add(label, family, alg, sect, f):
    fresh()
    ctx.sync()
    f()
    ctx.sync()
    rows.append((label, family, alg, sect))


# ---- general stride and reversed lanes: sdot / saxpy / srot at the configs[3] increments
if "strided" in GROUPS:
    for incx, incy in ((2, 2), (7, 7), (1, -1), (-1, -1), (1, 2), (7, 1)):
        sx, sy = sector_bytes(n, incx), sector_bytes(n, incy)
        fam = "unit(REV)" if (incx, incy) == (1, -1) else "unit" if (incx, incy) == (-1, -1) else "strided"
        add(f"sdot({incx},{incy})", "reduce_" + fam, 8 * n, sx + sy, lambda: ip.sdot_async(n, X, incx, Y, incy, out))
        add(f"saxpy({incx},{incy})", "map_" + fam, 12 * n, sx + 2 * sy, lambda: ip.saxpy_(n, 1.000123, X, incx, Y, incy))
        add(f"srot({incx},{incy})", "map_" + fam, 16 * n, 2 * sx + 2 * sy, lambda: ip.srot_(n, X, incx, Y, incy, c, s))
    add("sasum(7)", "reduce_strided", 4 * n, sector_bytes(n, 7), lambda: ip.sasum_async(n, X, 7, out))
    add("isamax(2)", "reduce_strided", 4 * n, sector_bytes(n, 2), lambda: ip.isamax_async(n, X, 2, out))

# ---- srotm's other functors (FLAG0 is in the headline capture)
if "rotm" in GROUPS:
    neg1 = lb.ModGivensRot(-1, 1.0, 1.0, 1.0, 0.9, -0.3, 0.2, 1.1)
    flag1 = lb.ModGivensRot(1, 1.0, 1.0, 1.0, 0.8, 0.0, 0.0, 0.7)
    add("srotm FLAGNEG1 (1,1)", "map_unit", 16 * n, 16 * n, lambda: ip.srotm_(neg1, n, X, 1, Y, 1))
    add("srotm FLAG1 (1,1)", "map_unit", 16 * n, 16 * n, lambda: ip.srotm_(flag1, n, X, 1, Y, 1))

# ---- x and y misaligned differently: the one-element-per-access shapes
if "misaligned" in GROUPS:
    xm, ym = X.view(1, n), Y.view(0, n)
    add("sdot misaligned(+1,+0)", "reduce_unit(VEC=1)", 8 * n, 8 * n, lambda: ip.sdot_async(n, xm, 1, ym, 1, out))
    add("saxpy misaligned(+1,+0)", "map_unit(VEC=1)", 12 * n, 12 * n, lambda: ip.saxpy_(n, 1.000123, xm, 1, ym, 1))
    add("scopy misaligned(+1,+0)", "map_unit(VEC=1)", 8 * n, 8 * n, lambda: ip.scopy_(n, xm, 1, ym, 1))
    add("sswap misaligned(+1,+0)", "map_unit(VEC=1)", 16 * n, 16 * n, lambda: ip.sswap_(n, xm, 1, ym, 1))
    xm, ym = X.view(5, n), Y.view(2, n)
    add("saxpy misaligned(+5,+2)", "map_unit(VEC=1)", 12 * n, 12 * n, lambda: ip.saxpy_(n, 1.000123, xm, 1, ym, 1))

# ---- read-only / write-only / mixed streams side by side (the DRAM turnaround question), unit stride
if "streams" in GROUPS:
    xu, yu = X.view(0, n), Y.view(0, n)
    add("sasum (pure read)", "reduce_unit", 4 * n, 4 * n, lambda: ip.sasum_async(n, xu, 1, out))
    add("sdot (pure read, two streams)", "reduce_unit", 8 * n, 8 * n, lambda: ip.sdot_async(n, xu, 1, yu, 1, out))
    add("fill (pure write)", "map_unit", 4 * n, 4 * n, lambda: xu.fill(1.5))
    add("sscal (1 read : 1 write, same lines)", "map_unit", 8 * n, 8 * n, lambda: ip.sscal_(n, 1.000123, xu, 1))
    add("scopy (1 read : 1 write)", "map_unit", 8 * n, 8 * n, lambda: ip.scopy_(n, xu, 1, yu, 1))
    add("saxpy (2 reads : 1 write)", "map_unit", 12 * n, 12 * n, lambda: ip.saxpy_(n, 1.000123, xu, 1, yu, 1))
    add("sswap (2 reads : 2 writes)", "map_unit", 16 * n, 16 * n, lambda: ip.sswap_(n, xu, 1, yu, 1))

# ---- every Double instance at the same bytes (n/2 doubles)
if "f64" in GROUPS:
    nd = n // 2
    Xd, Yd, outd = ctx.alloc(nd, np.float64), ctx.alloc(nd, np.float64), ctx.alloc(4, np.float64)
    flags = lb.drotmg(np.float64(1.5), np.float64(0.75), np.float64(2.0), np.float64(0.5))


    def addd(label, family, bytes_per, f):
        Xd.fill_hash(1); Yd.fill_hash(2)
        ctx.sync()
        f()
        ctx.sync()
        rows.append((label, family, bytes_per * nd, bytes_per * nd))


    addd("ddot", "reduce_unit", 16, lambda: ip.ddot_async(nd, Xd, 1, Yd, 1, outd))
    addd("dasum", "reduce_unit", 8, lambda: ip.dasum_async(nd, Xd, 1, outd))
    addd("dnrm2", "reduce_unit", 8, lambda: ip.dnrm2_async(nd, Xd, 1, outd))
    addd("idamax", "reduce_unit", 8, lambda: ip.idamax_async(nd, Xd, 1, outd))
    addd("daxpy", "map_unit", 24, lambda: ip.daxpy_(nd, 1.000123, Xd, 1, Yd, 1))
    addd("dscal", "map_unit", 16, lambda: ip.dscal_(nd, 1.000123, Xd, 1))
    addd("dcopy", "map_unit", 16, lambda: ip.dcopy_(nd, Xd, 1, Yd, 1))
    addd("dswap", "map_unit", 32, lambda: ip.dswap_(nd, Xd, 1, Yd, 1))
    addd("drot", "map_unit", 32, lambda: ip.drot_(nd, Xd, 1, Yd, 1, c, s))
    addd("drotm FLAG0", "map_unit", 32, lambda: ip.drotm_(flags, nd, Xd, 1, Yd, 1))
ctx.sync()
print("label,family,algorithmic_bytes,sector_bytes")
for r in rows:
    print(",".join(f'"{v}"' if isinstance(v, str) else str(v) for v in r))
ctx.close()

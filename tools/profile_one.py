"""A short program for ncu: one pass of each headline kernel at n = 2^28 (unit stride, in place).
    ncu --set full --clock-control none --import-source on -k regex:unit_kernel -c 10 -o gpurun_out/prof python tools/profile_one.py
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402
from lambda_blas_b200 import inplace as ip  # noqa: E402

n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 28
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = lb.Context(0)
x, y, out = ctx.alloc(n).fill_hash(1), ctx.alloc(n).fill_hash(2), ctx.alloc(4)
flags = lb.srotmg(1.5, 0.75, 2.0, 0.5)
cs, sn = math.cos(0.3), math.sin(0.3)
for _ in range(reps):
    ip.saxpy_(n, 1.000123, x, 1, y, 1)
    ip.sdot_async(n, x, 1, y, 1, out)
    ip.snrm2_async(n, x, 1, out)
    ip.sasum_async(n, x, 1, out)
    ip.isamax_async(n, x, 1, out)
    ip.sscal_(n, 1.000123, x, 1)
    ip.scopy_(n, x, 1, y, 1)
    ip.sswap_(n, x, 1, y, 1)
    ip.srot_(n, x, 1, y, 1, cs, sn)
    ip.srotm_(flags, n, x, 1, y, 1)
ctx.sync()
print("ok", out.to_host())
ctx.close()

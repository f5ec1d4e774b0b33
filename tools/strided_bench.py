"""BASELINE.json config 4: sdot / saxpy / srot with incx, incy in {1,2,7,-1} at n = 2^26.
Reports algorithmic GB/s (4 bytes per touched element per access; sector over-fetch is not credited)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402
from lambda_blas_b200 import inplace as ip  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 26)
# optional: "routine:variant:threads:bps" launch overrides, e.g. mapstrided:0:256:16 saxpy:0:128:64
overrides = [a.split(":") for a in sys.argv[2:]]
incs = [1, 2, 7, -1]
ctx = lb.Context(0)
for name, v, t, b in overrides:
    ctx.set_launch(name, int(v), int(t), int(b))
print("# overrides:", overrides)
span = 1 + (n - 1) * 7
x, y, out = ctx.alloc(span).fill_hash(1), ctx.alloc(span).fill_hash(2), ctx.alloc(4)
c, s = math.cos(0.3), math.sin(0.3)
B = {"sdot": 8, "saxpy": 12, "srot": 16, "sasum": 4, "isamax": 4, "snrm2": 4}
print("routine,incx,incy,us,GBps_algorithmic")
ROUTINES = os.environ.get("ROUTINES", "sdot,saxpy,srot").split(",")
for r in ROUTINES:
    for incx in incs:
        for incy in incs:
            if r in ("sasum", "isamax", "snrm2") and (incy != incs[0] or incx < 1):
                continue                      # one-vector reductions: positive incx only, once per incx
            f = {"sdot": lambda: ip.sdot_async(n, x, incx, y, incy, out),
                 "sasum": lambda: ip.sasum_async(n, x, incx, out), "isamax": lambda: ip.isamax_async(n, x, incx, out),
                 "snrm2": lambda: ip.snrm2_async(n, x, incx, out),
                 "saxpy": lambda: ip.saxpy_(n, 1.0, x, incx, y, incy),
                 "srot": lambda: ip.srot_(n, x, incx, y, incy, c, s)}[r]
            for _ in range(3):
                f()
            e0 = ctx.event().record()
            for _ in range(10):
                f()
            e1 = ctx.event().record()
            us = e0.elapsed_ms(e1) * 100
            print(f"{r},{incx},{incy},{us:.1f},{B[r] * n / us / 1e3:.1f}", flush=True)
            if r != "sdot":
                x.fill_hash(1); y.fill_hash(2)
ctx.close()

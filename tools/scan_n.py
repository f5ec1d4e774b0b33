"""Achieved GB/s of each routine (default launch shape) as n grows: separates the per-launch fixed
cost from the streaming rate.  time(n) ~ t0 + bytes(n)/rate."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402
from lambda_blas_b200 import inplace as ip  # noqa: E402

B = {"sdot": 8, "sasum": 4, "snrm2": 4, "isamax": 4, "saxpy": 12, "scopy": 8, "sswap": 16}
ctx = lb.Context(0)
nmax = 1 << 30
x, y, out = ctx.alloc(nmax).fill_hash(1), ctx.alloc(nmax).fill_hash(2), ctx.alloc(4)
print("routine,log2n,us,GBps")
for r in sys.argv[1].split(",") if len(sys.argv) > 1 else B:
    pts = []
    for lg in range(20, 31):
        n = 1 << lg
        f = {"sdot": lambda: ip.sdot_async(n, x, 1, y, 1, out), "sasum": lambda: ip.sasum_async(n, x, 1, out),
             "snrm2": lambda: ip.snrm2_async(n, x, 1, out), "isamax": lambda: ip.isamax_async(n, x, 1, out),
             "saxpy": lambda: ip.saxpy_(n, 1.0, x, 1, y, 1), "scopy": lambda: ip.scopy_(n, x, 1, y, 1),
             "sswap": lambda: ip.sswap_(n, x, 1, y, 1)}[r]
        for _ in range(3):
            f()
        iters = 20
        e0 = ctx.event().record()
        for _ in range(iters):
            f()
        e1 = ctx.event().record()
        us = e0.elapsed_ms(e1) / iters * 1e3
        pts.append((B[r] * n, us))
        print(f"{r},{lg},{us:.2f},{B[r] * n / us / 1e3:.1f}", flush=True)
    # least squares on the four largest sizes (all far beyond L2)
    p = pts[-4:]
    mx = sum(b for b, _ in p) / 4; my = sum(t for _, t in p) / 4
    slope = sum((b - mx) * (t - my) for b, t in p) / sum((b - mx) ** 2 for b, _ in p)
    print(f"# {r}: t0 = {my - slope * mx:.2f} us, streaming rate = {1 / slope / 1e3:.1f} GB/s", flush=True)
ctx.close()

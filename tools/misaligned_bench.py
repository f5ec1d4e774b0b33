"""Unit-stride routines on operands that are misaligned DIFFERENTLY (x starts `dx` elements, y `dy` elements past a
32-byte boundary): which kernel shape each case takes and what it moves.  n = 2^28 by default.
    python tools/misaligned_bench.py [log2n]  ->  CSV: routine,dx,dy,us,GBps,pct_of_8TBs
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402
from lambda_blas_b200 import inplace as ip  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
ctx = lb.Context(0)
# optional "routine:variant:threads:bps" launch overrides, e.g. mapmisaligned:24:512:1048576 redmisaligned:26:256:16
for name, v, t, b in (a.split(":") for a in sys.argv[2:]):
    ctx.set_launch(name, int(v), int(t), int(b))
print("# overrides:", sys.argv[2:])
PAIRS = [tuple(map(int, p.split("/"))) for p in os.environ.get("PAIRS", "0/0,1/1,1/0,0/1,3/0,4/0,4/1,5/2").split(",")]
X, Y, out = ctx.alloc(n + 16).fill_hash(1), ctx.alloc(n + 16).fill_hash(2), ctx.alloc(4)
c, s = math.cos(0.3), math.sin(0.3)
B = {"sdot": 8, "saxpy": 12, "scopy": 8, "sswap": 16, "srot": 16, "sasum": 4}
print("routine,dx,dy,us,GBps,pct_of_8TBs")
for r in os.environ.get("ROUTINES", "sdot,saxpy,scopy,sswap,srot,sasum").split(","):
    for dx, dy in PAIRS:
        x, y = X.view(dx, n), Y.view(dy, n)
        f = {"sdot": lambda: ip.sdot_async(n, x, 1, y, 1, out), "saxpy": lambda: ip.saxpy_(n, 1.0, x, 1, y, 1),
             "scopy": lambda: ip.scopy_(n, x, 1, y, 1), "sswap": lambda: ip.sswap_(n, x, 1, y, 1),
             "srot": lambda: ip.srot_(n, x, 1, y, 1, c, s), "sasum": lambda: ip.sasum_async(n, x, 1, out)}[r]
        for _ in range(3):
            f()
        e0 = ctx.event().record()
        for _ in range(10):
            f()
        e1 = ctx.event().record()
        us = e0.elapsed_ms(e1) * 100
        gbs = B[r] * n / us / 1e3
        print(f"{r},{dx},{dy},{us:.1f},{gbs:.1f},{gbs / 80:.1f}", flush=True)
        if r in ("saxpy", "srot"):
            X.fill_hash(1); Y.fill_hash(2)
ctx.close()

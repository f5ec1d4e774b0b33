"""Launch shapes of the unit-stride ELEMENTWISE kernels at mid sizes (2^20 .. 2^26): the defaults (one tile per block, 512 threads)
were tuned at n = 2^28, where the grid is tens of waves deep; at 2^22 a saxpy is 512 blocks on 296 resident slots.
SETS buffer sets are rotated so that no call finds its operands in the 126 MB L2; the walk is fixed forwards.
    python tools/midsize_maps.py [log2n ...] > gpurun_out/midsize_maps.csv
"""
import itertools
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402
from lambda_blas_b200 import inplace as ip  # noqa: E402

ctx = lb.Context(0)
ctx.set_traversal(0)
ONE = 1 << 20          # kOneTilePerBlock
B = {"saxpy": 12, "sscal": 8, "scopy": 8, "sswap": 16, "srot": 16}
NAME = {"saxpy": "axpy", "sscal": "scal", "scopy": "copy", "sswap": "swap", "srot": "rot"}
c, s = math.cos(0.3), math.sin(0.3)
print("routine,log2n,variant,threads,bps,us,GBps")
for lg in [int(a) for a in sys.argv[1:]] or [20, 22, 24, 25]:
    n = 1 << lg
    SETS = max(2, min(16, (1 << 30) // (4 * n)))
    xs = [ctx.alloc(n).fill_hash(1 + i) for i in range(SETS)]
    ys = [ctx.alloc(n).fill_hash(31 + i) for i in range(SETS)]
    calls = {"saxpy": lambda i: ip.saxpy_(n, 1e-3, xs[i], 1, ys[i], 1), "sscal": lambda i: ip.sscal_(n, 1.0001, xs[i], 1),
             "scopy": lambda i: ip.scopy_(n, xs[i], 1, ys[i], 1), "sswap": lambda i: ip.sswap_(n, xs[i], 1, ys[i], 1),
             "srot": lambda i: ip.srot_(n, xs[i], 1, ys[i], 1, c, s)}
    iters = 60
    for r in os.environ.get("ROUTINES", "saxpy,sscal,scopy,sswap,srot").split(","):
        rows = []
        shapes = itertools.chain([(-1, 0, 0)], itertools.product((16, 19, 20, 22, 17), (128, 256, 512), (ONE, 4, 8, 16, 32)))
        for v, t, b in shapes:
            ctx.set_launch(NAME[r], v, t, b)
            for i in range(2 * SETS):
                calls[r](i % SETS)
            e0 = ctx.event().record()
            for i in range(iters):
                calls[r](i % SETS)
            e1 = ctx.event().record()
            us = e0.elapsed_ms(e1) / iters * 1e3
            rows.append((us, v, t, b))
            print(f"{r},{lg},{v},{t},{b},{us:.2f},{B[r] * n / us / 1e3:.1f}", flush=True)
        ctx.set_launch(NAME[r], -1, 0, 0)
        best = sorted(rows)[:3]
        print(f"# {r} 2^{lg}: default {rows[0][0]:.2f} us; best " + "; ".join(f"v{v} t{t} b{b}: {us:.2f}" for us, v, t, b in best), flush=True)
    del xs, ys, calls
ctx.close()

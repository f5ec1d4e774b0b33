"""Launch shapes of the unit-stride reductions at MID sizes (the per-GPU share of n_global = 2^28 on 8 GPUs is 2^25): the
default shapes were tuned at n = 2^28, where a block runs ~28 tiles and the tail does not matter; at 2^25 it runs 3.5.
SETS buffer sets are rotated so that no call finds its operand in the 126 MB L2.
    python tools/midsize_sweep.py [log2n] > gpurun_out/midsize.csv
"""
import itertools
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402
from lambda_blas_b200 import inplace as ip  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 25
n = 1 << lg
SETS = max(2, min(8, (1 << 30) // (4 * n)))
ctx = lb.Context(0)
xs = [ctx.alloc(n).fill_hash(1 + i) for i in range(SETS)]
ys = [ctx.alloc(n).fill_hash(11 + i) for i in range(SETS)]
out = ctx.alloc(4)
B = {"sasum": 4, "sdot": 8, "snrm2": 4, "isamax": 4}
calls = {"sasum": lambda i: ip.sasum_async(n, xs[i], 1, out), "sdot": lambda i: ip.sdot_async(n, xs[i], 1, ys[i], 1, out),
         "snrm2": lambda i: ip.snrm2_async(n, xs[i], 1, out), "isamax": lambda i: ip.isamax_async(n, xs[i], 1, out)}
print("routine,log2n,variant,threads,bps,us,GBps")
iters = 40
for r in os.environ.get("ROUTINES", "sasum,sdot,snrm2,isamax").split(","):
    rows = []
    for v, t, b in itertools.chain([(-1, 0, 0)], itertools.product((3, 6, 4, 1, 7, 0), (256, 512), (3, 4, 6, 8, 12, 16, 24, 32, 48, 64))):
        if t * b > 512 * 64:
            continue
        ctx.set_launch(r, v, t, b)
        for i in range(2 * SETS):
            calls[r](i % SETS)
        e0 = ctx.event().record()
        for i in range(iters):
            calls[r](i % SETS)
        e1 = ctx.event().record()
        us = e0.elapsed_ms(e1) / iters * 1e3
        rows.append((us, v, t, b))
        print(f"{r},{lg},{v},{t},{b},{us:.2f},{B[r] * n / us / 1e3:.1f}", flush=True)
    ctx.set_launch(r, -1, 0, 0)
    best = sorted(rows)[:3]
    print(f"# {r} 2^{lg}: default {rows[0][0]:.2f} us; best " + "; ".join(f"v{v} t{t} b{b}: {us:.2f}" for us, v, t, b in best), flush=True)
ctx.close()

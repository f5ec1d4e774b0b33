"""Sweep launch shapes (variant x threads x blocks/SM) of the unit-stride kernels at n = 2^28 and
print achieved algorithmic GB/s (CUDA events on the library's stream).  Run on the GPU box:
    python tools/tune.py [--n 268435456] [--routines sdot,saxpy,...] > gpurun_out/tune.csv
"""
import argparse
import itertools
import math
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402
from lambda_blas_b200 import inplace as ip  # noqa: E402

BYTES = {"sdot": 8, "sdsdot": 8, "sasum": 4, "snrm2": 4, "isamax": 4, "saxpy": 12, "sscal": 8, "scopy": 8, "sswap": 16, "srot": 16, "srotm": 16, "fill": 4}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 28)
    ap.add_argument("--routines", default="sdot,saxpy,snrm2,scopy,sswap,isamax")
    ap.add_argument("--variants", default="0,1,2,3,4,5,6,7,9")
    ap.add_argument("--threads", default="128,256,512")
    ap.add_argument("--bps", default="0,-2,-4")   # 0: one resident wave; -k: k resident waves (grid = k * occupancy * SMs)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"], help="Float or Double instance (same bytes: n is halved for f64)")
    a = ap.parse_args()
    ctx = lb.Context(0)
    if os.environ.get("LBK_PDL"):
        lb._ffi.check(lb._ffi.lib().lbk_set_pdl(ctx.handle, int(os.environ["LBK_PDL"])), ctx.handle)
    import numpy as np
    dt = np.float64 if a.dtype == "f64" else np.float32
    esz = 8 if a.dtype == "f64" else 4
    n = a.n // (esz // 4)
    x, y = ctx.alloc(n, dt).fill_hash(1), ctx.alloc(n, dt).fill_hash(2)
    out = ctx.alloc(4, dt)
    flags = lb.srotmg(1.5, 0.75, 2.0, 0.5)
    c, s = math.cos(0.3), math.sin(0.3)
    calls = {
        "sdot": lambda: ip.sdot_async(n, x, 1, y, 1, out),
        "sasum": lambda: ip.sasum_async(n, x, 1, out),
        "sdsdot": lambda: ip.sdsdot_async(n, 0.5, x, 1, y, 1, out),
        "snrm2": lambda: ip.snrm2_async(n, x, 1, out),
        "isamax": lambda: ip.isamax_async(n, x, 1, out),
        "saxpy": lambda: ip.saxpy_(n, 1.000123, x, 1, y, 1),
        "sscal": lambda: ip.sscal_(n, 1.000123, x, 1),
        "scopy": lambda: ip.scopy_(n, x, 1, y, 1),
        "sswap": lambda: ip.sswap_(n, x, 1, y, 1),
        "srot": lambda: ip.srot_(n, x, 1, y, 1, c, s),
        "srotm": lambda: ip.srotm_(flags, n, x, 1, y, 1),
        "fill": lambda: y.fill(0.5),
    }
    print("routine,variant,threads,bps,ms,GBps")
    e0, e1 = ctx.event(), ctx.event()
    for r in a.routines.split(","):
        best = (0, None)
        for v, t, b in itertools.product([int(v) for v in a.variants.split(",")], [int(t) for t in a.threads.split(",")],
                                         [int(b) for b in a.bps.split(",")]):
            bps = b
            if b < 0:      # multiples of the occupancy-derived resident wave: resolved by trying a large explicit value
                bps = {128: 16, 256: 8, 512: 4}[t] * (-b)
            ctx.set_launch(r, v, t, bps)
            for _ in range(3):
                calls[r]()
            e0.record()
            for _ in range(a.iters):
                calls[r]()
            e1.record()
            ms = e0.elapsed_ms(e1) / a.iters
            gbs = BYTES[r] * (esz // 4) * n / ms / 1e6
            print(f"{r},{v},{t},{bps},{ms:.4f},{gbs:.1f}", flush=True)
            if gbs > best[0]:
                best = (gbs, (v, t, bps))
            # keep the data finite for the in-place routines
            if r in ("saxpy", "sscal", "srot", "srotm"):
                x.fill_hash(1); y.fill_hash(2)
        print(f"# best {r}: {best[0]:.1f} GB/s at variant,threads,bps = {best[1]}", flush=True)
        ctx.set_launch(r, -1, 0, 0)
    ctx.close()


if __name__ == "__main__":
    main()

"""Microseconds per kernel of recorded chains of GENERAL-STRIDE calls (inc 10, as in two thirds of test/Bench.hs), with the overlap
of programmatic dependent launch off and on; the chains are replayed as CUDA graphs so that the host is out of the picture.
    python tools/strided_chain.py  ->  CSV: chain,pdl,n,us_per_call
"""
import os
import sys
sys.path.insert(0, os.getcwd())
import lambda_blas_b200 as lb
from lambda_blas_b200 import inplace as ip
ctx = lb.Context(0)
print("chain,pdl,n,us_per_call")
for pdl in (0, 1, 0, 1):
    ctx.set_pdl(bool(pdl))
    for lg in (10, 14, 18):
        n = 1 << lg
        x, y, out = ctx.alloc(10 * n).fill_hash(1), ctx.alloc(10 * n).fill_hash(2), ctx.alloc(4)
        for name, f in (("saxpy inc 10 dependent", lambda: ip.saxpy_(n, 1e-3, x, 10, y, 10)),
                        ("sscal inc 10 + sdot inc 10", lambda: (ip.sscal_(n, 1.0001, x, 10), ip.sdot_async(n, x, 10, y, 10, out)))):
            with ctx.capture() as cap:
                for _ in range(100):
                    f()
            g = cap.graph
            for _ in range(3):
                g.launch()
            ctx.sync()
            a = ctx.event().record()
            for _ in range(20):
                g.launch()
            b = ctx.event().record()
            ctx.sync()
            print(f"{name},{pdl},{n},{a.elapsed_ms(b) * 1e3 / (20 * g.kernels):.2f}", flush=True)
ctx.close()

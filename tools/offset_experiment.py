"""Does the distance between the x stream and the y stream matter (DRAM bank / channel aliasing)?
scopy and saxpy at n = 2^28 with y placed `off` elements after a fixed base; x fixed.
    python tools/offset_experiment.py > gpurun_out/offset.csv
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402
from lambda_blas_b200 import inplace as ip  # noqa: E402

n = 1 << 28
ctx = lb.Context(0)
pad = 1 << 26
x = ctx.alloc(n).fill_hash(1)
ybig = ctx.alloc(n + pad).fill_hash(2)
print(f"# x at {x.ptr:#x}, ybig at {ybig.ptr:#x}, delta {(ybig.ptr - x.ptr) / 2**20:.3f} MiB")
print("routine,offset_bytes,ms,GBps")
for off in [0, 8, 64, 256, 1024, 4096, 16384, 65536, 1 << 18, 1 << 20, (1 << 20) + 4096, 1 << 22, (1 << 24) + (1 << 13), (1 << 25) + 24]:
    y = ybig.view(off, n)
    for name, f, b in (("scopy", lambda: ip.scopy_(n, x, 1, y, 1), 8), ("saxpy", lambda: ip.saxpy_(n, 1.0, x, 1, y, 1), 12)):
        for _ in range(3):
            f()
        e0 = ctx.event().record()
        for _ in range(20):
            f()
        e1 = ctx.event().record()
        ms = e0.elapsed_ms(e1) / 20
        print(f"{name},{off * 4},{ms:.4f},{b * n / ms / 1e6:.1f}", flush=True)
ctx.close()

"""DRAM counters over a STEADY stream of back-to-back launches (not one cold, serialised launch): ten launches of one routine
between cudaProfilerStart / cudaProfilerStop, for ncu's application-range replay:

    ncu --replay-mode app-range --metrics dram__bytes_read.sum,dram__bytes_write.sum,dram__cycles_active.sum,\
dram__cycles_active_read.sum,dram__cycles_active_write.sum,dram__cycles_elapsed.sum,gpu__time_duration.sum \
        --csv --log-file gpurun_out/steady.csv python tools/steady_counters.py

The question ("writers"): pure-read streams run at 93-94% of the nominal 8 TB/s, pure-write streams at 93%, every
read+write mix at 87-89% -- where do the DRAM cycles go?  One range per routine, in the order printed.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402
from lambda_blas_b200 import inplace as ip  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
reps = 10
ctx = lb.Context(0)
x, y, out = ctx.alloc(n).fill_hash(1), ctx.alloc(n).fill_hash(2), ctx.alloc(4)
rt = torch.cuda.cudart()
cases = [("sasum", 4, lambda: ip.sasum_async(n, x, 1, out)), ("sdot", 8, lambda: ip.sdot_async(n, x, 1, y, 1, out)),
         ("fill", 4, lambda: x.fill(1.5)), ("sscal", 8, lambda: ip.sscal_(n, 1.000123, x, 1)),
         ("scopy", 8, lambda: ip.scopy_(n, x, 1, y, 1)), ("saxpy", 12, lambda: ip.saxpy_(n, 1.000123, x, 1, y, 1)),
         ("sswap", 16, lambda: ip.sswap_(n, x, 1, y, 1))]
print("range,routine,launches,algorithmic_bytes")
for i, (name, b, f) in enumerate(cases):
    for _ in range(3):
        f()
    ctx.sync()
    rt.cudaProfilerStart()
    for _ in range(reps):
        f()
    ctx.sync()
    rt.cudaProfilerStop()
    print(f"{i},{name},{reps},{b * n * reps}", flush=True)
ctx.close()

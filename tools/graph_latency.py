"""Per-call latency of a chain of small calls: eager (programmatic dependent launch, one host call per kernel) against the same
chain recorded into a CUDA graph (lbk_capture_begin / lbk_capture_end) and replayed with one host call.
    python tools/graph_latency.py  ->  CSV: chain,log2n,calls,eager_us_per_call,graph_us_per_call,host_us_per_call_eager,host_us_per_replay
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402
from lambda_blas_b200 import inplace as ip  # noqa: E402

ctx = lb.Context(0)
CALLS, REPLAYS = 64, 20


def chains(n, x, y, z, out):
    return {
        "saxpy_ dependent (y += a x; 64 times)": lambda: ip.saxpy_(n, 1e-3, x, 1, y, 1),
        "sscal_ + sdot_async pairs": lambda: (ip.sscal_(n, 1.0001, x, 1), ip.sdot_async(n, x, 1, y, 1, out)),
        "independent (saxpy_ into y; sscal_ of z)": lambda: (ip.saxpy_(n, 1e-3, x, 1, y, 1), ip.sscal_(n, 1.0001, z, 1)),
        "srot_ + snrm2_async + isamax_async": lambda: (ip.srot_(n, x, 1, y, 1, 0.6, 0.8), ip.snrm2_async(n, x, 1, out), ip.isamax_async(n, y, 1, out.view(2, 2))),
    }


print("chain,log2n,calls,eager_us_per_call,graph_us_per_call,host_us_per_call_eager,host_us_per_replay")
for lg in (10, 14, 17, 20):
    n = 1 << lg
    x, y, z, out = ctx.alloc(n).fill_hash(1), ctx.alloc(n).fill_hash(2), ctx.alloc(n).fill_hash(3), ctx.alloc(8)
    for name, f in chains(n, x, y, z, out).items():
        before = ctx.launch_count()
        f()
        per = ctx.launch_count() - before
        reps = CALLS // per
        for _ in range(3):
            for _ in range(reps):
                f()
        ctx.sync()
        t0 = time.perf_counter()
        a = ctx.event().record()
        for _ in range(REPLAYS):
            for _ in range(reps):
                f()
        b = ctx.event().record()
        host_eager = (time.perf_counter() - t0) / (REPLAYS * reps * per) * 1e6
        ctx.sync()
        eager = a.elapsed_ms(b) * 1e3 / (REPLAYS * reps * per)
        with ctx.capture() as cap:
            for _ in range(reps):
                f()
        g = cap.graph
        for _ in range(3):
            g.launch()
        ctx.sync()
        t0 = time.perf_counter()
        a = ctx.event().record()
        for _ in range(REPLAYS):
            g.launch()
        b = ctx.event().record()
        host_graph = (time.perf_counter() - t0) / REPLAYS * 1e6
        ctx.sync()
        graph = a.elapsed_ms(b) * 1e3 / (REPLAYS * g.kernels)
        print(f"{name},{lg},{g.kernels},{eager:.2f},{graph:.2f},{host_eager:.2f},{host_graph:.1f}", flush=True)
ctx.close()

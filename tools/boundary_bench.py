"""What do the three kernel boundaries of the bench step cost?  Back-to-back runs of each routine alone, of each ordered pair and of
the whole triple (sdot, saxpy, snrm2 at n = 2^28, in place, PDL on): microseconds per sequence against the sum of its parts.
    python tools/boundary_bench.py [log2n]  ->  CSV: sequence,us_per_sequence,sum_of_parts_us,extra_us
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402
from lambda_blas_b200 import inplace as ip  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
ctx = lb.Context(0)
# optional "routine:variant:threads:bps" launch overrides and "pdl:0|1"
for a in sys.argv[2:]:
    f = a.split(":")
    if f[0] == "pdl":
        ctx.set_pdl(bool(int(f[1])))
    elif f[0] == "traversal":
        ctx.set_traversal(int(f[1]))
    else:
        ctx.set_launch(f[0], int(f[1]), int(f[2]), int(f[3]))
print("# overrides:", sys.argv[2:])
x, y, out = ctx.alloc(n).fill_hash(1), ctx.alloc(n).fill_hash(2), ctx.alloc(8)
F = {"sdot": lambda: ip.sdot_async(n, x, 1, y, 1, out), "saxpy": lambda: ip.saxpy_(n, 1e-6, x, 1, y, 1),
     "snrm2": lambda: ip.snrm2_async(n, x, 1, out.view(2, 2)), "sasum": lambda: ip.sasum_async(n, x, 1, out.view(4, 2)),
     "sscal": lambda: ip.sscal_(n, 1.0000001, y, 1)}
K = 60


def t(seq):
    for _ in range(5):
        for r in seq:
            F[r]()
    a = ctx.event().record()
    for _ in range(K):
        for r in seq:
            F[r]()
    b = ctx.event().record()
    ctx.sync()
    return a.elapsed_ms(b) * 1e3 / K


alone = {r: t([r]) for r in F}
print("sequence,us_per_sequence,sum_of_parts_us,extra_us")
for r in F:
    print(f"{r},{alone[r]:.1f},{alone[r]:.1f},0.0")
for seq in (["sdot", "saxpy"], ["saxpy", "snrm2"], ["snrm2", "sdot"], ["sdot", "saxpy", "snrm2"], ["saxpy", "sasum"], ["sscal", "snrm2"],
            ["sdot", "snrm2"], ["saxpy", "sdot"], ["sdot", "saxpy", "snrm2"]):
    us, parts = t(seq), sum(alone[r] for r in seq)
    print(f"{'+'.join(seq)},{us:.1f},{parts:.1f},{us - parts:.1f}", flush=True)
ctx.close()

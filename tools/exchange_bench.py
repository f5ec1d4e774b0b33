"""What the cross-GPU record exchange costs per reduction call, by size: back-to-back launches (rotating buffer sets, no L2 reuse) of
the same routine on SHARDED vectors (the exchange runs) against unsharded vectors of the same per-rank length on the same GPU (it does
not).  Run under torchrun:
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/exchange_bench.py > gpurun_out/exchange_nN.csv
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402
from lambda_blas_b200 import inplace as ip, sharded  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = sharded.attach_torch_distributed(lb.Context(local))
if len(sys.argv) > 1 and sys.argv[1] == "nccl":
    ctx.set_fused_exchange(False)
out = ctx.alloc(8)
B = {"sasum": 4, "snrm2": 4, "isamax": 4, "sdot": 8}
if rank == 0:
    print("routine,log2n_per_rank,ranks,us_sharded,us_local,exchange_us,GBps_per_gpu_sharded")
for lg in [int(v) for v in os.environ.get("SIZES", "22,23,24,25,26,27").split(",")]:
    n = 1 << lg
    SETS = max(2, min(8, (1 << 29) // (4 * n)))
    xs = [ctx.alloc_sharded(n * world).fill_hash(1 + i) for i in range(SETS)]
    ys = [ctx.alloc_sharded(n * world).fill_hash(11 + i) for i in range(SETS)]
    xl = [ctx.wrap(v.ptr, n, keepalive=v) for v in xs]          # the same memory as an unsharded vector of this rank
    yl = [ctx.wrap(v.ptr, n, keepalive=v) for v in ys]
    for r in B:
        def call(vx, vy, nn, i):
            if r == "sdot":
                ip.sdot_async(nn, vx[i], 1, vy[i], 1, out)
            else:
                getattr(ip, r + "_async")(nn, vx[i], 1, out)
        res = []
        for vx, vy, nn in ((xs, ys, n * world), (xl, yl, n)):
            for i in range(2 * SETS):
                call(vx, vy, nn, i % SETS)
            ctx.sync(); dist.barrier(); torch.cuda.synchronize()
            e0 = ctx.event().record()
            iters = 60
            for i in range(iters):
                call(vx, vy, nn, i % SETS)
            e1 = ctx.event().record()
            ctx.sync()
            t = torch.tensor([e0.elapsed_ms(e1) / iters * 1e3], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res.append(float(t.item()))
        if rank == 0:
            print(f"{r},{lg},{world},{res[0]:.2f},{res[1]:.2f},{res[0] - res[1]:.2f},{B[r] * n / res[0] / 1e3:.0f}", flush=True)
    del xl, yl, xs, ys
ctx.close()
dist.destroy_process_group()

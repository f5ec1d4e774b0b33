"""Sharded vectors over 2+ GPUs: one process per GPU, NCCL record exchange (run with gpurun --gpus 2).

Every rank must return the same bits; the result must match the oracle on the whole (unsharded) vector:
isamax exactly (ties and maxima at shard edges, NaN rule), sums within 2*n*eps*sum|.|, saxpy exactly.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import lambda_blas_b200 as lb
    from lambda_blas_b200 import inplace as ip, sharded
    from oracle import oracle
    ctx = sharded.attach_torch_distributed(lb.Context(rank))
    ok, notes = True, []
    fused_available = ctx.fused_exchange()
    cases = [(n, f, dt) for dt in (np.float32, np.float64) for f in ((True, False) if fused_available else (False,))
             for n in ([1, 2, 5, 1000, (1 << 20) + 3, 1 << 24] if dt == np.float32 else [1, 5, 1000, (1 << 20) + 3])]
    for n, fused, dt in cases:
        ctx.set_fused_exchange(fused)
        u = float(np.finfo(dt).eps) / 2          # unit round-off of the instance (Float: 2^-24, Double: 2^-53)
        bits = np.uint32 if dt == np.float32 else np.uint64
        x, y = ctx.alloc_sharded(n, dt).fill_hash(1), ctx.alloc_sharded(n, dt).fill_hash(2)
        hx, hy = lb.hash_fill_reference(n, 1, dtype=dt), lb.hash_fill_reference(n, 2, dtype=dt)
        lo, hi = sharded.shard_range(n, world, rank)
        assert len(x) == hi - lo and x.shard_offset == lo
        # plant a tie across the shard boundary and a maximum at the last element of shard 0
        if n >= 1000:
            edge = sharded.shard_range(n, world, 0)[1]
            for pos, val in ((edge - 1, 7.0), (edge, -7.0), (n - 1, 7.0)):
                hx[pos] = val
                if lo <= pos < hi:
                    x.view(pos - lo, 1).upload(np.array([val], dtype=dt)) if world == 1 else None
            # views of sharded vectors are refused by design: rewrite the shard through a wrapped handle
            shard = ctx.wrap(x.ptr, len(x), keepalive=x, dtype=dt)
            shard.upload(hx[lo:hi])
        prod = float(np.sum(np.abs(hx.astype(np.float64) * hy)))
        d = lb.sdot(n, x, 1, y, 1)
        ok &= abs(float(d) - float(oracle.sdot(n, hx, 1, hy, 1))) <= 2 * n * u * prod
        nr = lb.snrm2(n, x, 1)
        ok &= abs(float(nr) - float(oracle.snrm2(n, hx, 1))) <= 2 * n * u * float(oracle.snrm2(n, hx, 1))
        sa = lb.sasum(n, x, 1)
        ok &= abs(float(sa) - float(oracle.sasum(n, hx, 1))) <= 2 * n * u * float(np.sum(np.abs(hx)))
        im = lb.isamax(n, x, 1)
        ok &= im == oracle.isamax(n, hx, 1)
        if n >= 5:          # n smaller than the vector: only the leading ranks take part
            m = n - 2
            ok &= lb.isamax(m, x, 1) == oracle.isamax(m, hx, 1)
            ok &= abs(float(lb.sdot(m, x, 1, y, 1)) - float(oracle.sdot(m, hx, 1, hy, 1))) <= 2 * n * u * prod
        # every rank holds the same bits
        t = torch.tensor([float(d), float(nr), float(sa), float(im)], dtype=torch.float64, device="cuda")
        lo_t, hi_t = t.clone(), t.clone()
        dist.all_reduce(lo_t, op=dist.ReduceOp.MIN); dist.all_reduce(hi_t, op=dist.ReduceOp.MAX)
        ok &= bool(torch.equal(lo_t, hi_t))
        # elementwise: no communication, shard by shard
        ip.saxpy_(n, 1.000123, x, 1, y, 1)
        want_full = oracle.saxpy(n, 1.000123, hx, 1, hy, 1)
        ok &= np.array_equal(y.to_host().view(bits), want_full[lo:hi].view(bits))
        # equal positive increments on shards: logical element i sits at global position i*inc, so every pairing
        # stays on one rank (y has just been updated: hy follows)
        hy = want_full
        for inc in (2, 3, 7):
            m = (n - 1) // inc + 1
            prod = float(np.sum(np.abs(hx[::inc][:m].astype(np.float64) * hy[::inc][:m])))
            ok &= abs(float(lb.sdot(m, x, inc, y, inc)) - float(oracle.sdot(m, hx, inc, hy, inc))) <= 2 * m * u * prod
            ok &= abs(float(lb.snrm2(m, x, inc)) - float(oracle.snrm2(m, hx, inc))) <= 2 * m * u * float(oracle.snrm2(m, hx, inc))
            ok &= abs(float(lb.sasum(m, x, inc)) - float(oracle.sasum(m, hx, inc))) <= 2 * m * u * float(np.sum(np.abs(hx[::inc][:m])))
            ok &= lb.isamax(m, x, inc) == oracle.isamax(m, hx, inc)
            y2 = y.clone()
            ip.saxpy_(m, -0.75, x, inc, y2, inc)
            ok &= np.array_equal(y2.to_host().view(bits), oracle.saxpy(m, -0.75, hx, inc, hy, inc)[lo:hi].view(bits))
            x2, y3 = lb.srot(m, x, inc, y, inc, np.cos(0.3), np.sin(0.3))          # the pure form: new sharded vectors
            wx, wy = oracle.srot(m, hx, inc, hy, inc, np.cos(0.3), np.sin(0.3))
            ok &= np.array_equal(x2.to_host().view(bits), wx[lo:hi].view(bits)) and np.array_equal(y3.to_host().view(bits), wy[lo:hi].view(bits))
        # what would pair elements across devices is refused, not silently wrong
        for incx, incy in ((2, 3), (-2, -3), (1, -1), (0, 1)):
            try:
                lb.sdot(max(n // 4, 1), x, incx, y, incy)
                ok &= (world == 1)
            except lb.LbkError as e:
                ok &= e.status == 6
        notes.append((n, fused, np.dtype(dt).name, bool(ok)))
    notes.append(("fused_available", fused_available))
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put((int(t.item()), notes))
    ctx.close()
    dist.destroy_process_group()


def test_sharded_routines_match_oracle():
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    ok, notes = q.get(timeout=5)
    assert ok == 1, notes

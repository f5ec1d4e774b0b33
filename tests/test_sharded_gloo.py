"""The N > 1 path on CPU: world_size 2 over gloo.

What is exercised is the HOST logic a sharded reduction relies on -- the block partition
(lbk_shard_range), the per-rank record each GPU would contribute, the exchange in rank order and the
fold every rank performs (lbk_combine_records, the host statement of finish_records_kernel) -- with the
local kernels replaced by numpy.  The CUDA + NCCL path itself is tests/test_gpu_multi.py (-m gpu).
"""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _key(v):
    b = np.abs(np.float32(v)).view(np.uint32)
    return 0 if b > 0x7F800000 else int(b)


def _key64(v):
    b = np.abs(np.float64(v)).view(np.uint64)
    return 0 if b > 0x7FF0000000000000 else int(b)


def _local_record(kind, xs, ys, lo):
    """What a rank's reduction kernel leaves in its record for the shard xs (global offset lo)."""
    from lambda_blas_b200._ffi import Record
    r = Record((C.c_double * 3)(0, 0, 0), 0, -1, 0)
    if xs.size == 0:
        return r
    T = xs.dtype.type
    if kind == 0:
        r.f[0] = float(np.sum(xs * ys, dtype=T))
    elif kind == 1:
        r.f[1] = float(np.sum(xs ** 2, dtype=T))
    else:
        kf = _key64 if xs.dtype == np.float64 else _key
        keys = np.array([kf(v) for v in xs], dtype=np.uint64)
        i = int(np.argmax(keys))             # first maximum
        r.key, r.idx = int(keys[i]), lo + i
        if lo == 0 and np.isnan(xs[0]):
            r.key, r.idx = 2 ** 64 - 1, 0
    return r


def _worker(rank, world, port, seed, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lambda_blas_b200  # noqa: F401
    from lambda_blas_b200 import sharded
    from lambda_blas_b200._ffi import Record
    from oracle import oracle
    rng = np.random.default_rng(seed)           # same stream on every rank: the "global" vectors
    ok = True
    for n, dt in [(n, dt) for dt in (np.float32, np.float64) for n in [1, 2, 3, 101, 1000, 4097]]:
        eps = 2.0 ** -24 if dt == np.float32 else 2.0 ** -53
        x = rng.uniform(-1, 1, n).astype(dt)
        y = rng.uniform(-1, 1, n).astype(dt)
        if n >= 101:                             # a tie that straddles the shard boundary, and a NaN that must lose
            x[n // 2 - 1] = 5.0; x[n // 2 + 3] = -5.0; x[7] = np.nan
        lo, hi = sharded.shard_range(n, world, rank)
        # positive increments on shards (lbk_shard_span): logical element i sits at global position i*inc; this rank
        # reduces the ones inside its block, and the fold over ranks must give the oracle's answer on the whole vector
        for inc in (2, 3, 7):
            m = (n - 1) // inc + 1
            first, count, off = sharded.shard_span(n, world, rank, m, inc)
            xs, ys = x[lo:hi][off::inc][:count], y[lo:hi][off::inc][:count]
            ok &= np.array_equal(xs, x[::inc][first:first + count], equal_nan=True)
            mine = _local_record(2, xs, ys, first)
            buf = torch.frombuffer(bytearray(bytes(mine)), dtype=torch.uint8)
            gathered = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(gathered, buf)
            recs = [Record.from_buffer_copy(bytes(g.numpy().tobytes())) for g in gathered]
            ok &= sharded.combine_records(2, recs, dt == np.float64) == oracle.iamax(m, x, inc)
        for kind in (0, 1, 2):
            mine = _local_record(kind, x[lo:hi], y[lo:hi], lo)
            buf = torch.frombuffer(bytearray(bytes(mine)), dtype=torch.uint8)
            gathered = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(gathered, buf)
            recs = [Record.from_buffer_copy(bytes(g.numpy().tobytes())) for g in gathered]
            got = sharded.combine_records(kind, recs, dt == np.float64)
            if kind == 2:
                ok &= got == oracle.iamax(n, x, 1)
            elif kind == 0:
                xx = np.where(np.isnan(x), 0, x)
                want = float(np.dot(xx.astype(np.float64), y))
                rec2 = [Record.from_buffer_copy(bytes(r)) for r in recs]
                if not np.isnan(x).any():
                    ok &= abs(got - want) <= 2 * n * eps * float(np.sum(np.abs(xx.astype(np.float64) * y)))
                del rec2
            else:
                if not np.isnan(x).any():
                    ok &= abs(got - float(oracle.nrm2(n, x, 1))) <= 2 * n * eps * float(oracle.nrm2(n, x, 1))
    t = torch.tensor([1 if ok else 0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(int(t.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_reduction_host_logic(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 1234, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) == 1

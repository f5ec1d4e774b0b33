"""Multi-GPU plumbing: one process per GPU, vectors sharded by contiguous blocks.

The reference is single-process and single-threaded; sharding is new.  Elementwise
routines need no communication.  A reduction runs the local kernel, then exchanges ONE 48-byte
record per rank (NCCL all-gather over NVLink on the library's own stream) and every rank folds the
records in rank order, so all ranks return the same bits.  torch.distributed is used only to carry
the 128-byte NCCL id to the other ranks and for barriers around timed regions.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

from . import _ffi
from .types import Context


def shard_range(global_len: int, nranks: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of rank's block: the first (global_len % nranks) ranks hold one extra element."""
    lo, hi = C.c_int64(), C.c_int64()
    _ffi.check(_ffi.lib().lbk_shard_range(global_len, nranks, rank, C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def shard_span(global_len: int, nranks: int, rank: int, n: int, inc: int) -> Tuple[int, int, int]:
    """Of the logical elements i in [0, n) at global positions i*inc (inc >= 1): (first, count, offset) of those in
    rank's block, offset being where the first one sits inside the block."""
    a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
    _ffi.check(_ffi.lib().lbk_shard_span(global_len, nranks, rank, n, inc, C.byref(a), C.byref(b), C.byref(c)))
    return a.value, b.value, c.value


def attach_torch_distributed(ctx: Context) -> Context:
    """Give `ctx` an NCCL communicator spanning the initialised torch.distributed world."""
    import torch.distributed as dist
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return ctx
    box = [Context.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(world, rank, box[0])
    return ctx


def combine_records(kind: int, records, is_f64: bool = False) -> float | int:
    """Host statement of the cross-rank fold (lbk_combine_records); used by the CPU tests."""
    arr = (_ffi.Record * len(records))(*records)
    f, i = C.c_double(), C.c_int64()
    _ffi.check(_ffi.lib().lbk_combine_records(kind, int(is_f64), arr, len(records), C.byref(f), C.byref(i)))
    return i.value if kind == 2 else f.value

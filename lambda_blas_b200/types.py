"""Types of the path: the device vector added beside Numerical.BLAS.Types, and the reference's ADTs.

Reference: src/Numerical/BLAS/Types.hs:8 (GivensRot), :15-20 (ModGivensRot with constructors
FLAGNEG2 | FLAGNEG1{d1,d2,x1,h11,h12,h21,h22} | FLAG0{d1,d2,x1,h12,h21} | FLAG1{d1,d2,x1,h11,h22});
vectors in the reference are Data.Vector.Storable.Vector Float (Single.hs:67), i.e. a host buffer.
DVector is the device-resident representation the north star adds: an opaque handle to HBM with
explicit host transfer (to_device / from_device).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import NamedTuple, Optional

import numpy as np

from . import _ffi


def _is64(dtype) -> bool:
    return np.dtype(dtype) == np.float64


class Context:
    """One GPU, one stream, scratch for the reductions, and (optionally) the NCCL communicator."""

    def __init__(self, device: Optional[int] = None):
        L = _ffi.lib()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = C.c_void_p()
        _ffi.check(L.lbk_init(device, C.byref(h)))
        self._h = h
        self.device = device
        self.nranks, self.rank = 1, 0

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("context is closed")
        return self._h

    def close(self) -> None:
        """lbk_shutdown: vectors and events that are still alive keep the native context until they are freed."""
        if getattr(self, "_h", None) is not None:
            if getattr(self, "_borrowed", None) is None:      # a rank context goes with its multi-device context
                _ffi.lib().lbk_shutdown(self._h)
            self._h = None

    def set_exchange_timeout(self, milliseconds: int) -> None:
        """How long a sharded reduction waits for its peers' records before LBK_ERR_NCCL."""
        _ffi.check(_ffi.lib().lbk_set_exchange_timeout(self.handle, int(milliseconds)), self.handle)

    def sync(self) -> None:
        _ffi.check(_ffi.lib().lbk_sync(self.handle), self.handle)

    def sm_count(self) -> int:
        v = C.c_int()
        _ffi.check(_ffi.lib().lbk_sm_count(self.handle, C.byref(v)), self.handle)
        return v.value

    def launch_count(self) -> int:
        v = C.c_int64()
        _ffi.check(_ffi.lib().lbk_launch_count(self.handle, C.byref(v)), self.handle)
        return v.value

    def pdl_count(self) -> int:
        v = C.c_int64()
        _ffi.check(_ffi.lib().lbk_pdl_count(self.handle, C.byref(v)), self.handle)
        return v.value

    def shifted_count(self) -> int:
        """Launches that read a differently-misaligned x through aligned packs + a lane shift (lbk_shifted_count)."""
        v = C.c_int64()
        _ffi.check(_ffi.lib().lbk_shifted_count(self.handle, C.byref(v)), self.handle)
        return v.value

    def set_traversal(self, mode: int) -> None:
        """0: elementwise kernels always walk forwards; 1 (default): from the end the previous kernel left in the L2; 2: backwards."""
        _ffi.check(_ffi.lib().lbk_set_traversal(self.handle, int(mode)), self.handle)

    def set_pdl(self, enabled: bool) -> None:
        _ffi.check(_ffi.lib().lbk_set_pdl(self.handle, int(enabled)), self.handle)

    def set_launch(self, routine: str, variant: int = -1, threads: int = 0, blocks_per_sm: int = 0) -> None:
        _ffi.check(_ffi.lib().lbk_set_launch(self.handle, routine.encode(), variant, threads, blocks_per_sm), self.handle)

    def stream(self) -> int:
        return int(_ffi.lib().lbk_stream(self.handle) or 0)

    # ---- communicator (one process per GPU) ----
    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _ffi.check(_ffi.lib().lbk_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, nranks: int, rank: int, unique_id: bytes) -> None:
        assert len(unique_id) == 128
        _ffi.check(_ffi.lib().lbk_comm_init(self.handle, nranks, rank, unique_id), self.handle)
        self.nranks, self.rank = nranks, rank

    def fused_exchange(self) -> bool:
        """True when sharded reductions end inside the reduction kernel over NVLink peer memory."""
        v = C.c_int()
        _ffi.check(_ffi.lib().lbk_comm_transport(self.handle, C.byref(v)), self.handle)
        return bool(v.value)

    def set_fused_exchange(self, enabled: bool) -> None:
        _ffi.check(_ffi.lib().lbk_set_fused_exchange(self.handle, int(enabled)), self.handle)

    # ---- vectors (dtype: np.float32 = Vector Float, np.float64 = Vector Double) ----
    def alloc(self, n: int, dtype=np.float32) -> "DVector":
        h = C.c_void_p()
        fn = _ffi.lib().lbk_vec_alloc_f64 if _is64(dtype) else _ffi.lib().lbk_vec_alloc
        _ffi.check(fn(self.handle, n, C.byref(h)), self.handle)
        return DVector(self, h)

    def alloc_sharded(self, global_len: int, dtype=np.float32) -> "DVector":
        h = C.c_void_p()
        fn = _ffi.lib().lbk_vec_alloc_sharded_f64 if _is64(dtype) else _ffi.lib().lbk_vec_alloc_sharded
        _ffi.check(fn(self.handle, global_len, C.byref(h)), self.handle)
        return DVector(self, h)

    def wrap(self, device_ptr: int, n: int, keepalive=None, dtype=np.float32) -> "DVector":
        h = C.c_void_p()
        fn = _ffi.lib().lbk_vec_wrap_f64 if _is64(dtype) else _ffi.lib().lbk_vec_wrap
        _ffi.check(fn(self.handle, C.c_void_p(device_ptr), n, C.byref(h)), self.handle)
        v = DVector(self, h)
        v._keep = keepalive
        return v

    def to_device(self, host, dtype=None) -> "DVector":
        """Explicit host -> device transfer.  float64 input makes a double-precision vector, anything else single."""
        if dtype is None:
            dtype = np.float64 if getattr(host, "dtype", None) == np.float64 else np.float32
        a = np.ascontiguousarray(host, dtype=dtype)
        if a.ndim != 1:
            raise ValueError("vectors are one-dimensional")
        v = self.alloc(a.size, dtype)
        if a.size:
            _ffi.check(_ffi.lib().lbk_vec_upload(v.handle, a.ctypes.data_as(C.c_void_p), a.size), self.handle)
        return v

    # ---- CUDA graphs ----
    def capture(self) -> "_Capture":
        """`with ctx.capture() as cap: ...calls...` records the calls instead of running them; `cap.graph.launch()` replays them
        (lbk_capture_begin / lbk_capture_end: in-place and *_oop routines, *_async reductions, asynchronous copies)."""
        return _Capture(self)

    # ---- events ----
    def event(self) -> "Event":
        return Event(self)

    def wait_event(self, ev: "Event") -> None:
        """Work enqueued on this context from now on waits for `ev` (recorded on another context's stream)."""
        _ffi.check(_ffi.lib().lbk_wait_event(self.handle, ev._h), self.handle)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class MultiContext(Context):
    """ONE host thread, many GPUs (lbk_init_multi): rank r is devices[r].  `alloc_sharded` / `to_device_sharded` make
    vectors block-sharded over the ranks; every routine of `single` / `inplace` then fans out over the devices, and a
    reduction returns the scalar all ranks agree on.  `alloc` / `to_device` make unsharded vectors (whole on devices[0]).
    The same device may be listed more than once (the sharded path on a one-GPU box)."""

    def __init__(self, devices, threads: Optional[bool] = None):
        L = _ffi.lib()
        devices = [int(d) for d in devices]
        arr = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _ffi.check(L.lbk_init_multi(len(devices), arr, C.byref(h)))
        self._h = h
        self.device = devices[0]
        self.devices = devices
        self.nranks, self.rank = len(devices), 0
        if threads is not None:
            self.set_threads(threads)

    def set_threads(self, enabled: bool) -> None:
        """One launcher thread per device (True) or rank after rank from the calling thread (False)."""
        _ffi.check(_ffi.lib().lbk_set_multi_threads(self.handle, int(enabled)), self.handle)

    def rank_result(self, rank: int) -> bytes:
        """The 8 bytes rank `rank` itself holds for the most recent synchronous reduction (lbk_multi_result)."""
        buf = C.create_string_buffer(8)
        _ffi.check(_ffi.lib().lbk_multi_result(self.handle, rank, buf), self.handle)
        return buf.raw

    def rank_context(self, rank: int) -> "Context":
        """The borrowed one-device context of `rank` (for callers that drive the ranks themselves)."""
        cache = self.__dict__.setdefault("_ranks", {})
        if rank not in cache:
            h = C.c_void_p()
            _ffi.check(_ffi.lib().lbk_multi_rank(self.handle, rank, C.byref(h)), self.handle)
            c = Context.__new__(Context)
            c._h, c.device, c.nranks, c.rank, c._borrowed = h, self.devices[rank], self.nranks, rank, self
            cache[rank] = c
        return cache[rank]

    def to_device_sharded(self, host, dtype=None) -> "DVector":
        """Explicit host -> devices transfer of ONE host buffer into a vector sharded over the ranks."""
        if dtype is None:
            dtype = np.float64 if getattr(host, "dtype", None) == np.float64 else np.float32
        a = np.ascontiguousarray(host, dtype=dtype)
        if a.ndim != 1:
            raise ValueError("vectors are one-dimensional")
        v = self.alloc_sharded(a.size, dtype)
        if a.size:
            _ffi.check(_ffi.lib().lbk_vec_upload(v.handle, a.ctypes.data_as(C.c_void_p), a.size), self.handle)
        return v


class Graph:
    """A recorded sequence of calls (lbk_graph): `launch()` enqueues one replay on the context's stream."""

    def __init__(self, ctx: Context, handle):
        self._h, self.ctx = handle, ctx

    def launch(self) -> "Graph":
        _ffi.check(_ffi.lib().lbk_graph_launch(self._h), self.ctx.handle)
        return self

    @property
    def kernels(self) -> int:
        v = C.c_int64()
        _ffi.check(_ffi.lib().lbk_graph_kernels(self._h, C.byref(v)), self.ctx.handle)
        return v.value

    def __del__(self):
        try:
            if self._h is not None:
                _ffi.lib().lbk_graph_destroy(self._h)
        except Exception:
            pass
        self._h = None


class _Capture:
    def __init__(self, ctx: Context):
        self.ctx, self.graph = ctx, None

    def __enter__(self) -> "_Capture":
        _ffi.check(_ffi.lib().lbk_capture_begin(self.ctx.handle), self.ctx.handle)
        return self

    def __exit__(self, et, ev, tb):
        h = C.c_void_p()
        st = _ffi.lib().lbk_capture_end(self.ctx.handle, C.byref(h))
        if et is None:
            _ffi.check(st, self.ctx.handle)
            self.graph = Graph(self.ctx, h)
        elif st == 0:
            _ffi.lib().lbk_graph_destroy(h)       # the body raised: drop what was recorded
        return False


class Event:
    def __init__(self, ctx: Context):
        h = C.c_void_p()
        _ffi.check(_ffi.lib().lbk_event_create(ctx.handle, C.byref(h)), ctx.handle)
        self._h, self.ctx = h, ctx

    def record(self) -> "Event":
        _ffi.check(_ffi.lib().lbk_event_record(self._h), self.ctx.handle)
        return self

    def elapsed_ms(self, stop: "Event") -> float:
        ms = C.c_float()
        _ffi.check(_ffi.lib().lbk_event_elapsed_ms(self._h, stop._h, C.byref(ms)), self.ctx.handle)
        return ms.value

    def __del__(self):
        try:
            if self._h is not None:       # safe after ctx.close(): handles keep the native context alive
                _ffi.lib().lbk_event_destroy(self._h)
        except Exception:
            pass
        self._h = None


class DVector:
    """A device-resident fp32 vector (or this rank's contiguous shard of one)."""

    def __init__(self, ctx: Context, handle: C.c_void_p, parent: "Optional[DVector]" = None):
        self.ctx, self._h, self._parent, self._keep = ctx, handle, parent, None
        # a handle's geometry never changes: read it once (the pure routines ask for it on every call)
        L = _ffi.lib()
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _ffi.check(L.lbk_vec_len(handle, C.byref(a), C.byref(b), C.byref(c)), ctx.handle)
        self._len, self._glen, self._off = a.value, b.value, c.value
        self._ptr = int(L.lbk_vec_ptr(handle) or 0)
        self._dtype = np.dtype(np.float64 if L.lbk_vec_elem_size(handle) == 8 else np.float32)
        self._sharded = bool(L.lbk_vec_is_sharded(handle))

    @property
    def handle(self):
        return self._h

    def __len__(self) -> int:
        return self._len

    @property
    def global_len(self) -> int:
        return self._glen

    @property
    def shard_offset(self) -> int:
        return self._off

    @property
    def ptr(self) -> int:
        return self._ptr

    @property
    def dtype(self):
        return self._dtype

    @property
    def is_sharded(self) -> bool:
        return self._sharded

    def like(self) -> "DVector":
        """An uninitialised vector with this one's length, element type and shard layout (lbk_vec_alloc_like)."""
        h = C.c_void_p()
        _ffi.check(_ffi.lib().lbk_vec_alloc_like(self._h, C.byref(h)), self.ctx.handle)
        return DVector(self.ctx, h)

    def part(self, rank: int) -> "DVector":
        """Rank `rank`'s block of a multi-device context's vector, as a vector of that rank's context (borrowed)."""
        h = C.c_void_p()
        _ffi.check(_ffi.lib().lbk_vec_part(self._h, rank, C.byref(h)), self.ctx.handle)
        return DVector(self.ctx.rank_context(rank), h, parent=self)

    def clone(self) -> "DVector":
        h = C.c_void_p()
        _ffi.check(_ffi.lib().lbk_vec_clone(self._h, C.byref(h)), self.ctx.handle)
        return DVector(self.ctx, h)

    def view(self, offset: int, n: int) -> "DVector":
        h = C.c_void_p()
        _ffi.check(_ffi.lib().lbk_vec_view(self._h, offset, n, C.byref(h)), self.ctx.handle)
        return DVector(self.ctx, h, parent=self)

    def to_host(self) -> np.ndarray:
        """Explicit device -> host transfer (of this rank's shard; of the WHOLE vector on a multi-device context)."""
        n = len(self)
        out = np.empty(n, dtype=self.dtype)
        if n:
            _ffi.check(_ffi.lib().lbk_vec_download(self._h, out.ctypes.data_as(C.c_void_p), n), self.ctx.handle)
        return out

    def download(self, out: np.ndarray) -> np.ndarray:
        """Synchronous device -> host into an existing array (lbk_vec_download; pageable memory goes through staging lanes)."""
        assert out.dtype == self.dtype and out.flags.c_contiguous and out.size >= len(self)
        if len(self):
            _ffi.check(_ffi.lib().lbk_vec_download(self._h, out.ctypes.data_as(C.c_void_p), len(self)), self.ctx.handle)
        return out

    def upload(self, host) -> "DVector":
        a = np.ascontiguousarray(host, dtype=self.dtype)
        _ffi.check(_ffi.lib().lbk_vec_upload(self._h, a.ctypes.data_as(C.c_void_p), a.size), self.ctx.handle)
        return self

    def upload_async(self, host: np.ndarray) -> "DVector":
        """Enqueue host -> device (host should come from pinned_array); returns at once."""
        assert host.dtype == self.dtype and host.flags.c_contiguous
        _ffi.check(_ffi.lib().lbk_vec_upload_async(self._h, host.ctypes.data_as(C.c_void_p), host.size), self.ctx.handle)
        return self

    def download_async(self, out: np.ndarray) -> np.ndarray:
        """Enqueue device -> host into `out` (pinned); valid after ctx.sync()."""
        assert out.dtype == self.dtype and out.flags.c_contiguous
        _ffi.check(_ffi.lib().lbk_vec_download_async(self._h, out.ctypes.data_as(C.c_void_p), out.size), self.ctx.handle)
        return out

    def fill_hash(self, seed: int, lo: float = -1.0, hi: float = 1.0) -> "DVector":
        _ffi.check(_ffi.lib().lbk_vec_fill_hash(self._h, seed, lo, hi), self.ctx.handle)
        return self

    def fill(self, value: float) -> "DVector":
        _ffi.check(_ffi.lib().lbk_vec_fill(self._h, value), self.ctx.handle)
        return self

    def __del__(self):
        try:
            if self._h is not None:       # safe after ctx.close(): handles keep the native context alive
                _ffi.lib().lbk_vec_free(self._h)
        except Exception:
            pass
        self._h = None


class _Pinned:
    def __init__(self, nbytes: int):
        p = C.c_void_p()
        _ffi.check(_ffi.lib().lbk_host_alloc(nbytes, C.byref(p)))
        self.p, self.n = p, nbytes

    def __del__(self):
        try:
            _ffi.lib().lbk_host_free(self.p)
        except Exception:
            pass


def pinned_array(n: int, dtype=np.float32) -> np.ndarray:
    """A numpy array over page-locked host memory (lbk_host_alloc), for async transfers."""
    dtype = np.dtype(dtype)
    if n == 0:
        return np.empty(0, dtype=dtype)
    owner = _Pinned(n * dtype.itemsize)
    ctype = C.c_double if dtype == np.float64 else C.c_float
    arr = np.ctypeslib.as_array(C.cast(owner.p, C.POINTER(ctype)), shape=(n,))
    # keep the allocation alive as long as any view of the array is
    out = arr.view(_PinnedArray)
    out._owner = owner
    return out


class _PinnedArray(np.ndarray):
    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._owner = getattr(obj, "_owner", None)


class registered:
    """Context manager: page-lock a numpy array the caller already owns for the duration of the block
    (lbk_host_register / lbk_host_unregister), so uploads from it run at the rate of a pinned copy."""

    def __init__(self, array: np.ndarray):
        assert array.flags.c_contiguous
        self.a = array

    def __enter__(self):
        if self.a.nbytes:
            _ffi.check(_ffi.lib().lbk_host_register(self.a.ctypes.data_as(C.c_void_p), self.a.nbytes))
        return self.a

    def __exit__(self, *exc):
        if self.a.nbytes:
            _ffi.check(_ffi.lib().lbk_host_unregister(self.a.ctypes.data_as(C.c_void_p)))


def hash_fill_reference(n: int, seed: int, lo: float = -1.0, hi: float = 1.0, offset: int = 0, dtype=np.float32) -> np.ndarray:
    """numpy statement of lbk_vec_fill_hash: v[i] = lo + (hi-lo) * u24(splitmix64(seed*2^40 + offset+i)), in `dtype`."""
    R = np.dtype(dtype).type
    i = (np.arange(n, dtype=np.uint64) + np.uint64(offset)) + (np.uint64(seed) << np.uint64(40))
    with np.errstate(over="ignore"):
        z = i + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(R) * R(2.0 ** -24)
    span = R(R(hi) - R(lo))
    return (R(lo) + span * u).astype(R)


# ---- the reference's ADTs ---------------------------------------------------------------------------

class GivensRot(NamedTuple):
    """GIVENSROT (r, z, c, s)  -- Types.hs:8"""
    r: np.float32
    z: np.float32
    c: np.float32
    s: np.float32


@dataclass(frozen=True)
class ModGivensRot:
    """Types.hs:15-20.  `flag` names the constructor: -2 FLAGNEG2, -1 FLAGNEG1, 0 FLAG0, 1 FLAG1."""
    flag: int
    d1: float = np.float32(0)
    d2: float = np.float32(0)
    x1: float = np.float32(0)
    h11: float = np.float32(0)
    h12: float = np.float32(0)
    h21: float = np.float32(0)
    h22: float = np.float32(0)

    @property
    def constructor(self) -> str:
        return {-2: "FLAGNEG2", -1: "FLAGNEG1", 0: "FLAG0", 1: "FLAG1"}[self.flag]

    def sparam(self, dtype=np.float32) -> np.ndarray:
        """BLAS sparam [flag,h11,h21,h12,h22], filled the way test/Foreign.hs:276-280 pokes it."""
        if self.flag == -2:
            return np.array([-2, 1, 0, 0, 1], dtype=dtype)
        if self.flag == -1:
            return np.array([-1, self.h11, self.h21, self.h12, self.h22], dtype=dtype)
        if self.flag == 0:
            return np.array([0, 0, self.h21, self.h12, 0], dtype=dtype)
        return np.array([1, self.h11, 0, 0, self.h22], dtype=dtype)


FLAGNEG2 = ModGivensRot(-2)

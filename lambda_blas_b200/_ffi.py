"""ctypes binding of the C ABI in include/lbk.h (lambda_blas_b200/liblbk.so).

This is the same set of symbols the Haskell module hs/Numerical/BLAS/Device.hs imports with
`foreign import ccall`, in the same argument order.  There is NO fallback: if the CUDA library is
missing or no device is present, importing / creating a context raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LBK_LIBRARY: developer knob for A/B runs of differently built libraries (tools/); the default is the in-tree build
LIB_PATH = os.environ.get("LBK_LIBRARY") or os.path.join(_HERE, "liblbk.so")

STATUS_NAMES = {0: "LBK_OK", 1: "LBK_ERR_CUDA", 2: "LBK_ERR_NCCL", 3: "LBK_ERR_ARG", 4: "LBK_ERR_RANGE",
                5: "LBK_ERR_ALIAS", 6: "LBK_ERR_UNSUPPORTED", 7: "LBK_ERR_NOMEM"}


class LbkError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class Record(C.Structure):
    """lbk_record: what one rank contributes to a sharded reduction (48 bytes)."""
    _fields_ = [("f", C.c_double * 3), ("key", C.c_uint64), ("idx", C.c_int64), ("pad", C.c_int64)]


_i64, _f32, _f64, _vp, _int = C.c_int64, C.c_float, C.c_double, C.c_void_p, C.c_int
_pf32, _pf64, _pi64, _pvp, _pint = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_int)

# name -> (restype, argtypes); every symbol include/lbk.h declares
PROTOTYPES = {
    "lbk_version": (_int, []),
    "lbk_device_count": (_int, [_pint]),
    "lbk_init": (_int, [_int, _pvp]),
    "lbk_init_multi": (_int, [_int, _pint, _pvp]),
    "lbk_multi_rank": (_int, [_vp, _int, _pvp]),
    "lbk_multi_result": (_int, [_vp, _int, _vp]),
    "lbk_set_multi_threads": (_int, [_vp, _int]),
    "lbk_set_exchange_timeout": (_int, [_vp, _i64]),
    "lbk_shutdown": (_int, [_vp]),
    "lbk_sync": (_int, [_vp]),
    "lbk_last_error": (C.c_char_p, [_vp]),
    "lbk_stream": (_vp, [_vp]),
    "lbk_sm_count": (_int, [_vp, _pint]),
    "lbk_set_launch": (_int, [_vp, C.c_char_p, _int, _int, _int]),
    "lbk_launch_count": (_int, [_vp, _pi64]),
    "lbk_set_pdl": (_int, [_vp, _int]),
    "lbk_pdl_count": (_int, [_vp, _pi64]),
    "lbk_shifted_count": (_int, [_vp, _pi64]),
    "lbk_set_traversal": (_int, [_vp, _int]),
    "lbk_capture_begin": (_int, [_vp]),
    "lbk_capture_end": (_int, [_vp, _pvp]),
    "lbk_graph_launch": (_int, [_vp]),
    "lbk_graph_kernels": (_int, [_vp, _pi64]),
    "lbk_graph_destroy": (_int, [_vp]),
    "lbk_comm_unique_id": (_int, [C.c_char_p]),
    "lbk_comm_init": (_int, [_vp, _int, _int, C.c_char_p]),
    "lbk_comm_info": (_int, [_vp, _pint, _pint]),
    "lbk_comm_transport": (_int, [_vp, _pint]),
    "lbk_set_fused_exchange": (_int, [_vp, _int]),
    "lbk_shard_range": (_int, [_i64, _int, _int, _pi64, _pi64]),
    "lbk_shard_span": (_int, [_i64, _int, _int, _i64, _i64, _pi64, _pi64, _pi64]),
    "lbk_vec_alloc": (_int, [_vp, _i64, _pvp]),
    "lbk_vec_alloc_f64": (_int, [_vp, _i64, _pvp]),
    "lbk_vec_alloc_sharded": (_int, [_vp, _i64, _pvp]),
    "lbk_vec_alloc_sharded_f64": (_int, [_vp, _i64, _pvp]),
    "lbk_vec_wrap": (_int, [_vp, _vp, _i64, _pvp]),
    "lbk_vec_wrap_f64": (_int, [_vp, _vp, _i64, _pvp]),
    "lbk_vec_view": (_int, [_vp, _i64, _i64, _pvp]),
    "lbk_vec_alloc_like": (_int, [_vp, _pvp]),
    "lbk_vec_part": (_int, [_vp, _int, _pvp]),
    "lbk_vec_is_sharded": (_int, [_vp]),
    "lbk_vec_free": (_int, [_vp]),
    "lbk_vec_len": (_int, [_vp, _pi64, _pi64, _pi64]),
    "lbk_vec_elem_size": (_int, [_vp]),
    "lbk_vec_ptr": (_vp, [_vp]),
    "lbk_vec_upload": (_int, [_vp, _vp, _i64]),
    "lbk_vec_download": (_int, [_vp, _vp, _i64]),
    "lbk_vec_upload_async": (_int, [_vp, _vp, _i64]),
    "lbk_vec_download_async": (_int, [_vp, _vp, _i64]),
    "lbk_vec_clone": (_int, [_vp, _pvp]),
    "lbk_vec_fill_hash": (_int, [_vp, C.c_uint64, _f64, _f64]),
    "lbk_vec_fill": (_int, [_vp, _f64]),
    "lbk_host_alloc": (_int, [_i64, _pvp]),
    "lbk_host_free": (_int, [_vp]),
    "lbk_host_register": (_int, [_vp, _i64]),
    "lbk_host_unregister": (_int, [_vp]),
    "lbk_team_selftest": (_int, [_int, _int, _int, _int, _pi64, _pint]),
    "lbk_event_create": (_int, [_vp, _pvp]),
    "lbk_event_record": (_int, [_vp]),
    "lbk_event_elapsed_ms": (_int, [_vp, _vp, _pf32]),
    "lbk_wait_event": (_int, [_vp, _vp]),
    "lbk_event_destroy": (_int, [_vp]),
    "lbk_sdsdot": (_int, [_vp, _i64, _f32, _vp, _i64, _vp, _i64, _pf32]),
    "lbk_sdsdot_async": (_int, [_vp, _i64, _f32, _vp, _i64, _vp, _i64, _vp]),
    "lbk_srotg": (_int, [_f32, _f32, _pf32]),
    "lbk_drotg": (_int, [_f64, _f64, _pf64]),
    "lbk_srotmg": (_int, [_f32, _f32, _f32, _f32, _pf32, _pf32]),
    "lbk_drotmg": (_int, [_f64, _f64, _f64, _f64, _pf64, _pf64]),
    "lbk_combine_records": (_int, [_int, _int, C.POINTER(Record), _int, _pf64, _pi64]),
    "lbk_pdl_sim_create": (_int, [_pvp]),
    "lbk_pdl_sim_destroy": (_int, [_vp]),
    "lbk_pdl_sim_enable": (_int, [_vp, _int]),
    "lbk_pdl_sim_interrupt": (_int, [_vp]),
    "lbk_pdl_sim_op": (_int, [_vp, _int, _int, _pi64, _int, _pi64, _int, _pint, _pint]),
}
# the Float and Double instances of every vector routine: same shapes, s/d prefix, float/double scalars
for _p, _fx, _pfx, _ip in (("s", _f32, _pf32, "is"), ("d", _f64, _pf64, "id")):
    PROTOTYPES.update({
        f"lbk_{_p}dot": (_int, [_vp, _i64, _vp, _i64, _vp, _i64, _pfx]),
        f"lbk_{_p}asum": (_int, [_vp, _i64, _vp, _i64, _pfx]),
        f"lbk_{_p}nrm2": (_int, [_vp, _i64, _vp, _i64, _pfx]),
        f"lbk_{_ip}amax": (_int, [_vp, _i64, _vp, _i64, _pi64]),
        f"lbk_{_p}dot_async": (_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp]),
        f"lbk_{_p}asum_async": (_int, [_vp, _i64, _vp, _i64, _vp]),
        f"lbk_{_p}nrm2_async": (_int, [_vp, _i64, _vp, _i64, _vp]),
        f"lbk_{_ip}amax_async": (_int, [_vp, _i64, _vp, _i64, _vp]),
        f"lbk_{_p}axpy": (_int, [_vp, _i64, _fx, _vp, _i64, _vp, _i64]),
        f"lbk_{_p}scal": (_int, [_vp, _i64, _fx, _vp, _i64]),
        f"lbk_{_p}copy": (_int, [_vp, _i64, _vp, _i64, _vp, _i64]),
        f"lbk_{_p}swap": (_int, [_vp, _i64, _vp, _i64, _vp, _i64]),
        f"lbk_{_p}rot": (_int, [_vp, _i64, _vp, _i64, _vp, _i64, _fx, _fx]),
        f"lbk_{_p}rotm": (_int, [_vp, _i64, _vp, _i64, _vp, _i64, _pfx]),
        f"lbk_{_p}axpy_oop": (_int, [_vp, _i64, _fx, _vp, _i64, _vp, _i64, _vp]),
        f"lbk_{_p}scal_oop": (_int, [_vp, _i64, _fx, _vp, _i64, _vp]),
        f"lbk_{_p}copy_oop": (_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp]),
        f"lbk_{_p}swap_oop": (_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp]),
        f"lbk_{_p}rot_oop": (_int, [_vp, _i64, _vp, _i64, _vp, _i64, _fx, _fx, _vp, _vp]),
        f"lbk_{_p}rotm_oop": (_int, [_vp, _i64, _vp, _i64, _vp, _i64, _pfx, _vp, _vp]),
    })

_lib = None


def _try_build() -> None:
    """A checkout without the built library (it is git-ignored): compile it in tree if nvcc is here.
    This is the build step of __graft_entry__.build(), not a fallback -- without nvcc the import fails."""
    import shutil
    import subprocess
    if shutil.which("nvcc") and shutil.which("make"):
        subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-s"], check=False)


def lib() -> C.CDLL:
    """Load liblbk.so (built by lambda_blas_b200/csrc/Makefile or __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            _try_build()
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C lambda_blas_b200/csrc` (nvcc, sm_100a). "
                "lambda_blas_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)      # AttributeError here = header and library out of sync
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(status: int, ctx_handle=None) -> None:
    if status != 0:
        msg = lib().lbk_last_error(ctx_handle)
        raise LbkError(status, msg.decode() if msg else "")

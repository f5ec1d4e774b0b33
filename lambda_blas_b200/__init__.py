"""lambda_blas_b200 -- the single-precision Level-1 path of dlewissandy/lambda-blas on NVIDIA B200.

  single   : the reference's call surface (Numerical.BLAS.Single) over device vectors, pure results
  inplace  : the raw in-place / enqueue-only forms of the C ABI
  types    : Context, DVector, GivensRot, ModGivensRot
  sharded  : one process per GPU; torch.distributed carries the NCCL id, NCCL carries the records
  MultiContext : one process, one host thread, many GPUs (lbk_init_multi): the reference's call surface on sharded vectors

The compute path is hand-written CUDA for sm_100a in liblbk.so behind the C ABI of include/lbk.h.
Importing this package without that library raises ImportError -- there is no CPU fallback.
"""
from . import _ffi

_ffi.lib()          # fail loudly, at import, if the CUDA library is missing

from .types import Context, MultiContext, DVector, Event, registered, GivensRot, ModGivensRot, FLAGNEG2, hash_fill_reference, pinned_array  # noqa: E402
from . import single, inplace  # noqa: E402
from .single import (dot, sdsdot, asum, nrm2, iamax, scal, copy, swap, axpy, rot, rotm, rotg, rotmg,  # noqa: E402,F401
                     sdot, sasum, snrm2, isamax, sscal, scopy, sswap, saxpy, srot, srotm, srotg, srotmg,
                     ddot, dasum, dnrm2, idamax, dscal, dcopy, dswap, daxpy, drot, drotm, drotg, drotmg)
from ._ffi import LbkError  # noqa: E402,F401

__all__ = ["Context", "MultiContext", "registered", "DVector", "Event", "GivensRot", "ModGivensRot", "FLAGNEG2", "LbkError", "single", "inplace",
           "hash_fill_reference", "pinned_array"]

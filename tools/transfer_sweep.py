"""toDevice / fromDevice of one 2^28-element vector from PAGEABLE host memory through the library's staging lanes
(lbk_vec_upload / lbk_vec_download), for the lane count and chunk size in LBK_STAGE_LANES / LBK_STAGE_CHUNK_MB."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lambda_blas_b200 as lb  # noqa: E402

n = 1 << 28
ctx = lb.Context(0)
x = ctx.alloc(n)
h = np.ones(n, dtype=np.float32)
o = np.zeros(n, dtype=np.float32)
for name, f in (("up", lambda: x.upload(h)), ("down", lambda: x.download(o))):
    f()
    t0 = time.perf_counter()
    for _ in range(3):
        f()
    dt = (time.perf_counter() - t0) / 3
    print(f"lanes={os.environ.get('LBK_STAGE_LANES', '4')},chunk_mb={os.environ.get('LBK_STAGE_CHUNK_MB', '8')},{name},{4 * n / dt / 1e9:.1f} GB/s", flush=True)
ctx.close()

"""Generates tests/golden/ref_golden_mid.json: the REFERENCE'S OWN SOURCE (evaluated by oracle/hs_eval.py, like
make_ref_golden.py) at sizes where the CUDA path runs its multi-block kernels, every unit-stride shape (forward, reversed
lanes, misaligned views) and the general-stride ones: n = 4099 and 20011.

Inputs are not stored: x = counter_hash(seed 1), y = counter_hash(seed 2) (tests/gen.py; lbk_vec_fill_hash makes the same values
on the device).  Vector results are stored as the SHA-256 of their bytes -- the comparison is bit for bit either way, and the
inputs hold no NaN -- reductions as bit patterns.

Run from the repo root IN THE BUILD CONTAINER (needs /root/reference; a few minutes):  python tests/golden/make_ref_golden_mid.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hs_eval  # noqa: E402
from tests.gen import counter_hash, span  # noqa: E402

PAIRS = [(1, 1), (1, -1), (-1, 1), (-1, -1), (2, 3), (-3, 1), (7, 7), (1, 0), (0, 1)]
A = 1.000123


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(v, dt):
    return int(np.asarray(v, dtype=dt).reshape(1).view(np.uint32 if dt == np.float32 else np.uint64)[0])


def main():
    R = hs_eval.Reference()
    out = {"about": "outputs of the reference's own source (src/Numerical/BLAS/{Types,Util,Single}.hs) evaluated by oracle/hs_eval.py at "
                    "n = 4099 / 20011; inputs = tests.gen.counter_hash(span, seed 1 / 2); vectors as sha256 of their bytes; "
                    "see tests/golden/make_ref_golden_mid.py", "a": A, "c": float(np.cos(0.3)), "s": float(np.sin(0.3)),
           "rotmg_args": [1.5, 0.75, 2.0, 0.5], "cases": []}
    for dt in (np.float32, np.float64):
        mgr = R.rotmg(*out["rotmg_args"], ty=dt)
        for n in (4099, 20011):
            for ix, iy in PAIRS:
                if n > 5000 and (ix, iy) not in ((1, 1), (1, -1), (-1, -1), (2, 3), (7, 7)):
                    continue
                x, y = counter_hash(span(n, ix, 2), 1, dtype=dt), counter_hash(span(n, iy, 1), 2, dtype=dt)
                a, c, s = dt(A), dt(np.cos(0.3)), dt(np.sin(0.3))
                case = {"dtype": np.dtype(dt).name, "n": n, "incx": ix, "incy": iy, "mgr": mgr[0],
                        "sdot": bits(R.dot(n, x, ix, y, iy), dt), "scopy": sha(R.copy(n, x, ix, y, iy)),
                        "saxpy": sha(R.axpy(n, a, x, ix, y, iy))}
                if dt == np.float32:
                    case["sdsdot"] = bits(R.sdsdot(n, a, x, ix, y, iy), dt)
                if ix != 0 and iy != 0:
                    case["sswap"] = [sha(v) for v in R.swap(n, x, ix, y, iy)]
                    case["srot"] = [sha(v) for v in R.rot(n, x, ix, y, iy, c, s)]
                    case["srotm"] = [sha(v) for v in R.rotm(mgr, n, x, ix, y, iy)]
                if ix > 0:
                    case.update({"sasum": bits(R.asum(n, x, ix), dt), "snrm2": bits(R.nrm2(n, x, ix), dt), "isamax": R.iamax(n, x, ix),
                                 "sscal": sha(R.scal(n, a, x, ix))})
                out["cases"].append(case)
                print(case["dtype"], n, ix, iy, flush=True)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden_mid.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print(path, len(out["cases"]), "cases", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

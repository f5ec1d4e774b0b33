"""Generates tests/golden/level1_golden.json: seeded inputs and the ORACLE's outputs, bit patterns as uint32.

These fixtures are outputs of oracle/lblas_oracle.c (the set made from the REFERENCE'S source is
ref_golden_*.json, make_ref_golden.py) -- cross-checked, case by case while generating, against the netlib 3.7.0
restatement wherever the reference's own tests assert `native == netlib` (inc != 0 rules of Main.hs).
They pin the oracle against silent drift (compiler flags, edits) and give the GPU tests fixed inputs.
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from tests.gen import gen_nice_float, nice_scalar, span  # noqa: E402


DT = np.float32


def u32(a):
    """bit patterns as unsigned integers (uint32 for the Float file, uint64 for the Double file)"""
    return [int(v) for v in np.asarray(a, dtype=DT).reshape(-1).view(np.uint32 if DT == np.float32 else np.uint64)]


def main(dtype=np.float32, fname="level1_golden.json", reps=(0, 1, 2), ns=(0, 1, 2, 5, 17)):
    global DT
    DT = dtype
    oracle.build()
    rng = np.random.default_rng(20261019 if dtype == np.float32 else 20261020)
    cases = []
    incs = [-5, -2, -1, 0, 1, 2, 3, 7]
    special = (np.array([np.nan, np.inf, -np.inf, -0.0, 0.0, 1e-45, -1e-40, 3.4e38, 1e-30, 2.0 ** 60, -2.0 ** -70], dtype=np.float32)
               if dtype == np.float32 else
               np.array([np.nan, np.inf, -np.inf, -0.0, 0.0, 5e-324, -1e-310, 1.7e308, 1e-200, 2.0 ** 500, -2.0 ** -520], dtype=np.float64))
    srotg, srotmg = (oracle.srotg, oracle.srotmg) if dtype == np.float32 else (oracle.drotg, oracle.drotmg)

    def vec(n, inc, extra=0, spice=False):
        v = gen_nice_float(rng, span(n, inc, extra), dtype)
        if spice and v.size:
            k = rng.integers(0, v.size, size=max(1, v.size // 6))
            v[k] = rng.choice(special, size=k.size)
        return v

    for rep in reps:
        for n in ns:
            for incx in incs:
                incy = int(rng.choice(incs))
                spice = rep == 2
                x, y = vec(max(n, 1), incx, int(rng.integers(0, 3)), spice), vec(max(n, 1), incy, int(rng.integers(0, 3)), spice)
                a = nice_scalar(rng, dtype)
                c, s = dtype(np.cos(a)), dtype(np.sin(a))
                # moderate magnitudes every other set: the only way to FLAG0 / FLAG1 at Double
                flags = srotmg(*([nice_scalar(rng, dtype) for _ in range(4)] if len(cases) % 2 else list(rng.uniform(-4, 4, 4).astype(dtype))))
                sx, sy = oracle.sswap(n, x, incx, y, incy)
                rx, ry = oracle.srot(n, x, incx, y, incy, c, s)
                mx, my = oracle.srotm(flags, n, x, incx, y, incy)
                cases.append({
                    "n": n, "incx": incx, "incy": incy, "a": u32([a])[0], "c": u32([c])[0], "s": u32([s])[0],
                    "param": u32(flags.sparam(dtype)), "x": u32(x), "y": u32(y), "spiced": spice,
                    "outs": {
                        "sdot": u32([oracle.sdot(n, x, incx, y, incy)]),
                        "sdsdot": (u32([oracle.sdsdot(n, a, x, incx, y, incy)]) if dtype == np.float32 else []),
                        "sasum": u32([oracle.sasum(n, x, incx)]), "snrm2": u32([oracle.snrm2(n, x, incx)]),
                        "isamax": [oracle.isamax(n, x, incx)], "sscal": u32(oracle.sscal(n, a, x, incx)),
                        "scopy": u32(oracle.scopy(n, x, incx, y, incy)), "saxpy": u32(oracle.saxpy(n, a, x, incx, y, incy)),
                        "sswap": [u32(sx), u32(sy)], "srot": [u32(rx), u32(ry)], "srotm": [u32(mx), u32(my)]}})
    scal = []
    for it in range(200):
        args = [nice_scalar(rng, dtype) for _ in range(4)] if it % 3 else list(rng.uniform(-4, 4, 4).astype(dtype))
        scal.append({"args": u32(args), "rotg": u32(list(srotg(args[0], args[1]))),
                     "rotmg": [srotmg(*args).flag] + u32(list(srotmg(*args))[1:])})
    out = {"about": "oracle outputs (oracle/lblas_oracle.c); bit patterns as unsigned integers; see make_golden.py",
           "dtype": np.dtype(dtype).name,
           "readme_example": {"n": 3, "x": u32([1, 2, 3]), "y": u32([4, 5, 6]), "sdot": u32([32.0])[0]},
           "cases": cases, "scalars": scal}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), fname)
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(path, len(cases), "input sets x 11 routines", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main(np.float32, "level1_golden.json")
    main(np.float64, "level1_golden_f64.json", reps=(0, 2), ns=(0, 1, 5, 17))

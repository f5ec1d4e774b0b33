"""Generates tests/golden/ref_golden_f32.json and ref_golden_f64.json: seeded inputs and the outputs of the
REFERENCE'S OWN SOURCE (/root/reference/src/Numerical/BLAS/{Types,Util,Single}.hs) evaluated by oracle/hs_eval.py.

This is the one place where the reference itself produces the expected values: there is no GHC in this image, so
its three hot-path modules are evaluated from their text (see the header of oracle/hs_eval.py for what is the
reference and what is a restatement of base / vector primitives).  The fixtures travel to the GPU box, where
/root/reference does not exist; the C oracle (tests/test_ref_golden.py) and the CUDA path
(tests/test_gpu_ref_golden.py) are checked against them.

Domains follow the reference's properties (test/Main.hs:28-52): n in 1..100, increments in [-5..5] (every pair
of them appears at least once), genNiceFloat / genNiceDouble elements, vectors of 1+(n-1)|inc| elements plus 0..2
spare ones; on top of that n = 0, and "spiced" sets with NaN / Inf / -0 / denormal / huge / tiny elements that the
reference's generators never produce.  Bit patterns are stored as unsigned integers.

Run from the repo root IN THE BUILD CONTAINER (needs /root/reference):  python tests/golden/make_ref_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import hs_eval  # noqa: E402
from tests.gen import gen_nice_float, nice_scalar, span  # noqa: E402

INCS = list(range(-5, 6))


def bits(a, dt):
    return [int(v) for v in np.asarray(a, dtype=dt).reshape(-1).view(np.uint32 if dt == np.float32 else np.uint64)]


def sparam(mgr, dt):
    """BLAS sparam layout [flag,h11,h21,h12,h22] as test/Foreign.hs:276-280 fills it"""
    con, f = mgr
    if con == "FLAGNEG2":
        return np.array([-2, 1, 0, 0, 1], dtype=dt)
    if con == "FLAGNEG1":
        return np.array([-1, f["h11"], f["h21"], f["h12"], f["h22"]], dtype=dt)
    if con == "FLAG0":
        return np.array([0, 0, f["h21"], f["h12"], 0], dtype=dt)
    return np.array([1, f["h11"], 0, 0, f["h22"]], dtype=dt)


def main(dt, fname, seed):
    R = hs_eval.Reference()
    rng = np.random.default_rng(seed)
    special = (np.array([np.nan, np.inf, -np.inf, -0.0, 0.0, 1e-45, -1e-40, 3.4e38, 1e-30, 2.0 ** 60, -2.0 ** -70], dtype=np.float32)
               if dt == np.float32 else
               np.array([np.nan, np.inf, -np.inf, -0.0, 0.0, 5e-324, -1e-310, 1.7e308, 1e-200, 2.0 ** 500, -2.0 ** -520], dtype=np.float64))

    def vec(n, inc, spice):
        v = gen_nice_float(rng, span(max(n, 1), inc, int(rng.integers(0, 3))), dt)
        if spice:
            k = rng.integers(0, v.size, size=max(1, v.size // 5))
            v[k] = rng.choice(special, size=k.size)
        return v

    plan = []       # (n, incx, incy, spiced)
    for ix in INCS:                                     # every increment pair of the reference's domain, small n
        for iy in INCS:
            plan.append((int(rng.integers(1, 7)), ix, iy, False))
    for _ in range(48):                                 # the long end of `choose (1,100)`
        plan.append((int(rng.integers(7, 101)), int(rng.choice(INCS)), int(rng.choice(INCS)), False))
    for ix in INCS:                                     # n = 0
        plan.append((0, ix, int(rng.choice(INCS)), False))
    for _ in range(44):                                 # values outside the reference's generators
        plan.append((int(rng.integers(1, 12)), int(rng.choice(INCS)), int(rng.choice(INCS)), True))

    cases = []
    seen_flags = {}
    for n, ix, iy, spice in plan:
        x, y = vec(n, ix, spice), vec(n, iy, spice)
        a = nice_scalar(rng, dt)
        c, s = dt(np.cos(float(a))), dt(np.sin(float(a)))
        # rotmg arguments: the reference's generator every other case, moderate magnitudes otherwise (the only
        # way to reach FLAG0 / FLAG1 at Double)
        margs = [nice_scalar(rng, dt) for _ in range(4)] if len(cases) % 2 else list(rng.uniform(-4, 4, 4).astype(dt))
        mgr = R.rotmg(*margs, ty=dt)
        seen_flags[mgr[0]] = seen_flags.get(mgr[0], 0) + 1
        sx, sy = R.swap(n, x, ix, y, iy)
        rx, ry = R.rot(n, x, ix, y, iy, c, s)
        mx, my = R.rotm(mgr, n, x, ix, y, iy)
        cases.append({
            "n": n, "incx": ix, "incy": iy, "a": bits([a], dt)[0], "c": bits([c], dt)[0], "s": bits([s], dt)[0],
            "mgr": mgr[0], "param": bits(sparam(mgr, dt), dt), "x": bits(x, dt), "y": bits(y, dt), "spiced": spice,
            "outs": {
                "sdot": bits([R.dot(n, x, ix, y, iy)], dt),
                "sdsdot": bits([R.sdsdot(n, a, x, ix, y, iy)], dt) if dt == np.float32 else [],
                "sasum": bits([R.asum(n, x, ix)], dt), "snrm2": bits([R.nrm2(n, x, ix)], dt),
                "isamax": [R.iamax(n, x, ix)], "sscal": bits(R.scal(n, a, x, ix), dt),
                "scopy": bits(R.copy(n, x, ix, y, iy), dt), "saxpy": bits(R.axpy(n, a, x, ix, y, iy), dt),
                "sswap": [bits(sx, dt), bits(sy, dt)], "srot": [bits(rx, dt), bits(ry, dt)],
                "srotm": [bits(mx, dt), bits(my, dt)]}})

    scal = []
    tiny, huge = (1e-12, 1e12) if dt == np.float32 else (1e-40, 1e40)
    for it in range(320):
        k = it % 4
        if k == 0:
            args = list(rng.uniform(-4, 4, 4).astype(dt))
        elif k == 1:
            args = [nice_scalar(rng, dt) for _ in range(4)]
        elif k == 2:      # d1 / d2 far outside (rgamsq, gamsq): the rescaling loops of checkscale (Single.hs:358-401)
            args = [dt(rng.uniform(0.5, 2) * rng.choice([tiny, huge, 1.0])), dt(rng.uniform(0.5, 2) * rng.choice([tiny, huge, 1.0])),
                    dt(rng.uniform(-2, 2)), dt(rng.uniform(-2, 2))]
        else:             # zeros and signs: the early exits (sd1 < 0, sp2 == 0, sq2 < 0)
            args = [dt(rng.choice([-1.0, 0.0, 1.0, 2.0])), dt(rng.choice([-1.0, 0.0, 1.0, 0.5])),
                    dt(rng.choice([0.0, 1.0, -3.0])), dt(rng.choice([0.0, 2.0, -1.0]))]
        g = R.rotg(args[0], args[1], ty=dt)
        con, f = R.rotmg(*args, ty=dt)
        seen_flags[con] = seen_flags.get(con, 0) + 1
        scal.append({"args": bits(args, dt), "rotg": bits(list(g), dt), "rotmg": [con] + bits([f[k] for k in hs_eval.MGR_FIELDS[con]], dt)})

    out = {"about": "outputs of the reference's own source (src/Numerical/BLAS/{Types,Util,Single}.hs of dlewissandy/lambda-blas) "
                    "evaluated by oracle/hs_eval.py; bit patterns as unsigned integers; see tests/golden/make_ref_golden.py",
           "dtype": np.dtype(dt).name, "seed": seed, "constructors_seen": seen_flags,
           "readme_example": {"n": 3, "x": bits([1, 2, 3], dt), "y": bits([4, 5, 6], dt),
                              "sdot": bits([R.dot(3, np.array([1, 2, 3], dt), 1, np.array([4, 5, 6], dt), 1)], dt)[0]},
           "cases": cases, "scalars": scal}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), fname)
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(path, len(cases), "input sets x 11 routines,", len(scal), "scalar cases,", os.path.getsize(path) // 1024, "KiB", seen_flags)


if __name__ == "__main__":
    main(np.float32, "ref_golden_f32.json", 20261020)
    main(np.float64, "ref_golden_f64.json", 20261021)

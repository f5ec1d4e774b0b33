"""Input generators re-stating the reference's QuickCheck generators (test/Gen.hs).

genNiceFloat (Gen.hs:39-44): frequency 10:40:40 of 0, +everyday, -everyday;
genEveryday (Gen.hs:130-131, 208-214): uniform on the real interval (2^-24, 2^24-1).
Seeded (numpy Generator) so every run sees the same cases -- the reference's are unseeded.
"""
import numpy as np

EVERYDAY_LO = np.float32(2.0 ** -24)
EVERYDAY_HI = np.float32(2.0 ** 24 - 1)


def everyday_range(dtype=np.float32):
    """Gen.hs:208-214: (2^-p, 2^p - 1) with p = floatDigits: 24 for Float, 53 for Double."""
    p = 24 if np.dtype(dtype) == np.float32 else 53
    return 2.0 ** -p, 2.0 ** p - 1


def gen_everyday(rng: np.random.Generator, size, dtype=np.float32) -> np.ndarray:
    lo, hi = everyday_range(dtype)
    return rng.uniform(lo, hi, size=size).astype(dtype)


def gen_nice_float(rng: np.random.Generator, size, dtype=np.float32) -> np.ndarray:
    """genNiceFloat / genNiceDouble (Gen.hs:39-52)."""
    kind = rng.choice(3, size=size, p=[10 / 90, 40 / 90, 40 / 90])
    v = gen_everyday(rng, size, dtype)
    v = np.where(kind == 0, 0, np.where(kind == 1, v, -v))
    return v.astype(dtype)


def nice_scalar(rng, dtype=np.float32):
    return gen_nice_float(rng, 1, dtype)[0]


def gen_unit(rng: np.random.Generator, size, dtype=np.float32) -> np.ndarray:
    """U(-1,1): the large-n bench distribution."""
    return rng.uniform(-1.0, 1.0, size=size).astype(dtype)


def span(n: int, inc: int, extra: int = 0) -> int:
    """Vector length the reference's tests allocate: extra + 1 + (n-1)|inc| (Main.hs:72,122,251)."""
    return extra + 1 + max(n - 1, 0) * abs(inc)


def counter_hash(n: int, seed: int, lo: float = -1.0, hi: float = 1.0, dtype=np.float32) -> np.ndarray:
    """v[i] = lo + (hi-lo) * (top 24 bits of splitmix64(seed * 2^40 + i)) * 2^-24 in `dtype`: the values lbk_vec_fill_hash puts on
    the device (inputs from a counter-based hash, so that fixtures need only carry the seed)."""
    R = np.dtype(dtype).type
    z = np.arange(n, dtype=np.uint64) + (np.uint64(seed) << np.uint64(40))
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = (z >> np.uint64(40)).astype(R) * R(2.0 ** -24)
    return (R(lo) + R(R(hi) - R(lo)) * u).astype(R)

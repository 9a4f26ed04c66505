"""KoalaBear helpers for host-side test/bench data (numpy, uint64 intermediates).

p = 2^31 - 2^24 + 1.  In-memory form at the C ABI is Montgomery with R = 2^32, the representation
of `KoalaBear` inside the reference's `RowMajorMatrix<Val>` (the in-tree C++ reinterpret-casts
it to `kb31_t`, crates/core/machine/cpp/extern.cpp:12; constants kb31_t.hpp:27-34).
"""
import numpy as np

P = 0x7F000001
R_MOD_P = (1 << 32) % P          # 0x1fffffe
R_INV = pow(R_MOD_P, P - 2, P)


def _chunked(fn, x, chunk=1 << 24):
    x = np.ascontiguousarray(x)
    if x.size <= chunk:
        return fn(x)
    out = np.empty(x.shape, dtype=np.uint32)
    flat_in, flat_out = x.reshape(-1), out.reshape(-1)
    for i in range(0, x.size, chunk):
        flat_out[i:i + chunk] = fn(flat_in[i:i + chunk])
    return out


def to_monty(x) -> np.ndarray:
    return _chunked(lambda a: ((a.astype(np.uint64) << np.uint64(32)) % np.uint64(P)).astype(np.uint32), np.asarray(x))


def from_monty(x) -> np.ndarray:
    return _chunked(lambda a: ((a.astype(np.uint64) * np.uint64(R_INV)) % np.uint64(P)).astype(np.uint32), np.asarray(x))


def mul(a, b) -> np.ndarray:
    return ((np.asarray(a, dtype=np.uint64) * np.asarray(b, dtype=np.uint64)) % np.uint64(P)).astype(np.uint32)


def add(a, b) -> np.ndarray:
    return ((np.asarray(a, dtype=np.uint64) + np.asarray(b, dtype=np.uint64)) % np.uint64(P)).astype(np.uint32)


def sub(a, b) -> np.ndarray:
    return ((np.asarray(a, dtype=np.uint64) + np.uint64(P) - np.asarray(b, dtype=np.uint64)) % np.uint64(P)).astype(np.uint32)


def random_elements(rng: np.random.Generator, shape) -> np.ndarray:
    return rng.integers(0, P, size=shape, dtype=np.uint32)


# SepticDigest::zero(): CURVE_CUMULATIVE_SUM_START_X ++ _Y (canonical), crates/stark/src/septic_digest.rs:9-14.
# The global cumulative sum of every Local-scope chip and of a program without initial memory.
SEPTIC_DIGEST_ZERO = np.array([637514027, 1595065213, 1998064738, 72333738, 1211544370, 822986770, 1518535784, 1604177449, 90440090, 259343427, 140470264, 1162099742, 941559812, 1064053343], dtype=np.uint32)

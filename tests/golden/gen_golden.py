#!/usr/bin/env python3
"""Generates tests/golden/ref_vectors.json from the REFERENCE's own C++ headers
(crates/core/machine/include/kb31_t.hpp, crates/recursion/core/include/poseidon2*.hpp) compiled
into oracle/_ref/libzkref.so by `make -C oracle ref`.  Needs /root/reference, so it only runs in
the build container; the JSON it writes is committed and is what the tests read everywhere else.

Also records the reference's in-tree known-answer test for this path
(examples/poseidon2/host/src/main.rs:33-37)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_ffi as o  # noqa: E402

P = 0x7F000001
ref = o.ref_lib()
assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
rng = np.random.Generator(np.random.PCG64(2025))
a = rng.integers(0, P, 64, dtype=np.uint32)
b = rng.integers(0, P, 64, dtype=np.uint32)
edge = np.array([0, 1, 2, P - 1, P - 2, 1 << 24, (1 << 24) - 1, 0x7EFFFFFF], dtype=np.uint32)
a[:8], b[:8] = edge, edge[::-1]
vec = {
    "source": "oracle/_ref/libzkref.so = reference headers kb31_t.hpp + poseidon2_skinny.hpp (event_to_row)",
    "a": a.tolist(), "b": b.tolist(),
    "mul": [ref.ref_kb31_mul(int(x), int(y)) for x, y in zip(a, b)],
    "add": [ref.ref_kb31_add(int(x), int(y)) for x, y in zip(a, b)],
    "sub": [ref.ref_kb31_sub(int(x), int(y)) for x, y in zip(a, b)],
    "inv": [ref.ref_kb31_inv(int(x)) if x else 0 for x in a],
    "to_monty": [ref.ref_kb31_to_monty(int(x)) for x in a],
    "poseidon2": [],
    "kat_poseidon2_hash_of_1000_ones": "ae45b14fe23b9f584c76c67d4d9ef6635a27b553a7114427584cc87ba8919866",
}
states = [np.arange(16, dtype=np.uint32), np.zeros(16, np.uint32), np.full(16, P - 1, np.uint32)]
states += [rng.integers(0, P, 16, dtype=np.uint32) for _ in range(13)]
for s in states:
    buf = (C.c_uint32 * 16)(*[int(x) for x in s])
    ref.ref_poseidon2_permute(buf)
    vec["poseidon2"].append({"in": s.tolist(), "out": list(buf)})
json.dump(vec, open(os.path.join(os.path.dirname(__file__), "ref_vectors.json"), "w"), indent=0)
print("wrote ref_vectors.json")

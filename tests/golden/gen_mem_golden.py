"""Writes tests/golden/mem_access.json: MemoryReadCols as the REFERENCE'S OWN C++ fills them
(crates/core/machine/include/memory.hpp populate_read, compiled into oracle/_ref/libzkref_core.so by `make -C oracle
ref`), for edge and seeded random MemoryReadRecords.  Run in the build container (needs /root/reference)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_ffi as o  # noqa: E402

recs = [(0, 1, 1, 1, 0), (0xFFFFFFFF, 1, 1 << 23, 1, 0), (0x01020304, 5, 100, 5, 99), (0x80000000, 5, 100, 4, 7_000_000),
        (7, 1 << 16, 3, 1, 0xFFFFFF), (0xDEADBEEF, 2, (1 << 24) - 1, 2, 0), (1, 9, 0x10001, 9, 0), (2, 9, 0x10000, 9, 0)]
rng = np.random.default_rng(0x3E3)
for _ in range(56):
    sh = int(rng.integers(1, 1 << 16))
    same = bool(rng.integers(0, 2))
    psh = sh if same else int(rng.integers(0, sh))
    ts = int(rng.integers(1, 1 << 24))
    pts = int(rng.integers(0, ts)) if same else int(rng.integers(0, 1 << 24))
    recs.append((int(rng.integers(0, 1 << 32)), sh, ts, psh, pts))
cols = []
for r in recs:
    c = o.ref_mem_access(*r)
    assert c is not None, "oracle/_ref/libzkref_core.so is missing: make -C oracle ref"
    cols.append([int(x) for x in c])
json.dump({"source": "crates/core/machine/include/memory.hpp populate_read via oracle/_ref/libzkref_core.so",
           "record_fields": ["value", "shard", "timestamp", "prev_shard", "prev_timestamp"],
           "records": recs, "read_cols": cols}, open(os.path.join(ROOT, "tests", "golden", "mem_access.json"), "w"))
print(len(recs), "records")

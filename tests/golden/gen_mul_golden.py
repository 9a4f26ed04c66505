"""Writes tests/golden/mul_rows.json: MulChip rows as the REFERENCE'S OWN C++ fills them (crates/core/machine/include/mul.hpp,
compiled into oracle/_ref/libzkref_core.so by `make -C oracle ref`) for the reference's own test event
(crates/core/machine/src/alu/mul/mod.rs:546-574) and 95 edge / seeded random CompAluEvent records.  Run in the build
container (needs /root/reference)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_ffi as o  # noqa: E402
from ziren_b200 import tracegen as tg  # noqa: E402

ref_event = [5, 790405, 1017624, 1017628, 3, 241306, 1298966409, 274417, 3776743705, 241306, 5, 790409, 3431, 5, 790387, 1]
ev = np.concatenate([np.array([ref_event], np.uint32), tg.synthetic_mul_events(95, seed=7)])
rows = o.ref_mul_rows(ev)
assert rows is not None, "oracle/_ref/libzkref_core.so is missing: make -C oracle ref"
json.dump({"source": "crates/core/machine/include/mul.hpp event_to_row via oracle/_ref/libzkref_core.so", "width": int(rows.shape[1]),
           "events": ev.tolist(), "rows": rows.tolist()}, open(os.path.join(ROOT, "tests", "golden", "mul_rows.json"), "w"))
print(ev.shape, rows.shape)

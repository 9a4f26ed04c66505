"""Writes tests/golden/cpu_rows.json: Cpu rows as the REFERENCE'S OWN C++ fills them (crates/core/machine/include/cpu.hpp,
compiled into oracle/_ref/libzkref_core.so by `make -C oracle ref`) for the reference's own test event
(crates/core/machine/src/cpu/trace.rs:283-306: an ADD whose b / c records hold another value than b / c, shard 0) and 119
seeded records covering every opcode.  Run in the build container (needs /root/reference)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_ffi as o  # noqa: E402
from ziren_b200 import tracegen as tg  # noqa: E402

# clk, pc, next_pc, next_next_pc, a, b, c, hi, flags (hi Some | a Write | b Read | c Read | imm_c), ADD | 29 << 8 | shard 0 << 16,
# op_b, op_c, a_record Write(5, 1, 2, 1, 1, 1), b_record Read(5, 0, 1, 0, 0), c_record Read(5, 0, 2, 0, 0)
ref_event = [0, 0, 1, 2, 5, 10, 15, 1, 1 | (2 << 1) | (1 << 3) | (1 << 4) | (1 << 6), 0 | (29 << 8), 0, 1,
             5, 1, 2, 1, 1, 1, 5, 0, 1, 0, 0, 5, 0, 2, 0, 0]
ev = np.concatenate([np.array([ref_event], np.uint32), tg.synthetic_cpu_events(119, seed=7)])
rows = o.ref_cpu_rows(ev)
assert rows is not None, "oracle/_ref/libzkref_core.so is missing: make -C oracle ref"
json.dump({"source": "crates/core/machine/include/cpu.hpp event_to_row via oracle/_ref/libzkref_core.so", "width": int(rows.shape[1]),
           "events": ev.tolist(), "rows": rows.tolist()}, open(os.path.join(ROOT, "tests", "golden", "cpu_rows.json"), "w"))
print(ev.shape, rows.shape)

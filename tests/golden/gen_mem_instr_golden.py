"""Writes tests/golden/mem_instr_rows.json: MemoryInstrs rows as the REFERENCE'S OWN C++ fills them
(crates/core/machine/include/memory_instrs.hpp, compiled into oracle/_ref/libzkref_core.so by `make -C oracle ref`) for 112
seeded MemInstrEvent records: every load and store opcode at every byte offset, negative bytes and halfwords, register-range
addresses, previous accesses in this and in an earlier shard.  Run in the build container (needs /root/reference)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_ffi as o  # noqa: E402
from ziren_b200 import tracegen as tg  # noqa: E402

ev = tg.synthetic_mem_instr_events(112, seed=11)
rows = o.ref_mem_instr_rows(ev)
assert rows is not None, "oracle/_ref/libzkref_core.so is missing: make -C oracle ref"
json.dump({"source": "crates/core/machine/include/memory_instrs.hpp event_to_row via oracle/_ref/libzkref_core.so",
           "width": int(rows.shape[1]), "events": ev.tolist(), "rows": rows.tolist()},
          open(os.path.join(ROOT, "tests", "golden", "mem_instr_rows.json"), "w"))
print(ev.shape, rows.shape)

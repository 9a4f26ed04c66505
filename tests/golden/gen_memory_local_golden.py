"""Writes tests/golden/memory_local_entries.json: SingleMemoryLocal entries as the REFERENCE'S OWN C++ fills them
(crates/core/machine/include/memory_local.hpp, compiled into oracle/_ref/libzkref_core.so by `make -C oracle ref`) for 90
seeded MemoryLocalEvent records (registers, untouched memory, earlier-shard initial accesses).  Run in the build container
(needs /root/reference)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_ffi as o  # noqa: E402
from ziren_b200 import tracegen as tg  # noqa: E402

ev = tg.synthetic_memory_local_events(90, seed=4)
entries = o.ref_memory_local_entries(ev)
assert entries is not None, "oracle/_ref/libzkref_core.so is missing: make -C oracle ref"
json.dump({"source": "crates/core/machine/include/memory_local.hpp event_to_row via oracle/_ref/libzkref_core.so",
           "entry_width": int(entries.shape[1]), "events": ev.tolist(), "entries": entries.tolist()},
          open(os.path.join(ROOT, "tests", "golden", "memory_local_entries.json"), "w"))
print(ev.shape, entries.shape)

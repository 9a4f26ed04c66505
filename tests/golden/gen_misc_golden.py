"""Writes tests/golden/misc_rows.json: MiscInstrs rows as the REFERENCE'S OWN C++ fills them
(crates/core/machine/include/misc_instrs.hpp, compiled into oracle/_ref/libzkref_core.so by `make -C oracle ref`) for 112
seeded well-formed MiscEvent records (SEXT byte / halfword, EXT, INS, MADDU, MSUBU, MADD, MSUB, TEQ).  Run in the build
container (needs /root/reference)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_ffi as o  # noqa: E402
from ziren_b200 import tracegen as tg  # noqa: E402

ev = tg.synthetic_misc_events(112, seed=13)
rows = o.ref_misc_rows(ev)
assert rows is not None, "oracle/_ref/libzkref_core.so is missing: make -C oracle ref"
json.dump({"source": "crates/core/machine/include/misc_instrs.hpp event_to_row via oracle/_ref/libzkref_core.so",
           "width": int(rows.shape[1]), "events": ev.tolist(), "rows": rows.tolist()},
          open(os.path.join(ROOT, "tests", "golden", "misc_rows.json"), "w"))
print(ev.shape, rows.shape)

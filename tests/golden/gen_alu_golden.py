#!/usr/bin/env python3
"""Writes tests/golden/alu_rows.json: trace rows of the core ALU chips produced by the REFERENCE'S
OWN C++ row fillers (crates/core/machine/include/{add_sub,bitwise,lt,shift_left,shift_right,
clo_clz}.hpp, compiled where they lie into oracle/_ref/libzkref_core.so by `make -C oracle ref`).
Run in the build container (needs /root/reference); the JSON travels to the GPU box.
Per chip: 96 events (every 4th pair of edge operands, equal and one-bit-apart operands, seeded
random ones) as 7 words {pc, next_pc, opcode, hi, a, b, c} and the rows as Montgomery words."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_ffi as o  # noqa: E402
from ziren_b200 import tracegen as tg  # noqa: E402

out = {"source": "reference C++ event_to_row via oracle/_ref/libzkref_core.so (tests/golden/gen_alu_golden.py)", "chips": {}}
for chip in tg.ALU_CHIPS:
    full = tg.synthetic_events(chip, 4096, seed=7)
    m = len(tg.EDGE_OPERANDS) ** 2
    pick = np.concatenate([np.arange(0, m, 4)[:56], np.arange(m, m + 8), np.arange(m + 256, m + 264), np.arange(3000, 3024)])
    ev = full[pick]
    rows = o.ref_alu_rows(chip, ev)
    assert rows is not None, "build oracle/_ref first (make -C oracle ref)"
    out["chips"][chip] = {"width": int(rows.shape[1]), "events": ev.tolist(), "rows": rows.tolist()}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "alu_rows.json")
with open(path, "w") as f:
    json.dump(out, f, separators=(",", ":"))
print("wrote", path, os.path.getsize(path), "bytes")

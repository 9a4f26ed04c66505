"""Writes tests/golden/more_rows.json: rows of DivRem, SyscallCore, SyscallPrecompile, SyscallInstrs and the per-event columns
of MemoryGlobalInit / MemoryGlobalFinalize as the REFERENCE'S OWN C++ fills them (crates/core/machine/include/{div_rem,syscall,
syscall_instrs,memory_global}.hpp, compiled into oracle/_ref/libzkref_core.so by `make -C oracle ref`) for seeded well-formed
event records.  DivRem: only events with c != 0 and not INT_MIN / -1 - div_rem.hpp differs from the Rust generate_trace there
(INT32_MAX instead of u32::MAX as the quotient for c = 0, max(1, |c|) in abs_c, a native INT_MIN / -1).  Run in the build
container (needs /root/reference)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_ffi as o  # noqa: E402
from ziren_b200 import tracegen as tg  # noqa: E402


def twin_domain(ev):
    """DivRem events on which div_rem.hpp and divrem/mod.rs agree."""
    b, c = ev[:, 7], ev[:, 8]
    return ev[(c != 0) & ~((b == 0x80000000) & (c == 0xFFFFFFFF))]


def cases():
    yield "DivRem", 16, twin_domain(tg.synthetic_div_rem_events(160, seed=21))[:96]
    yield "SyscallCore", 14, tg.synthetic_syscall_events(64, seed=22, kind="core")
    yield "SyscallPrecompile", 14, tg.synthetic_syscall_events(64, seed=23, kind="precompile")
    yield "SyscallInstrs", 14, tg.synthetic_syscall_events(96, seed=24, kind="instrs")
    ev = tg.synthetic_memory_global_events(48, seed=25)
    yield "MemoryGlobalInit", 4, ev[np.argsort(ev[:, 0])]


if __name__ == "__main__":
    out = {"source": "crates/core/machine/include/{div_rem,syscall,syscall_instrs,memory_global}.hpp event_to_row via oracle/_ref/libzkref_core.so"}
    for chip, words, ev in cases():
        rows = o.ref_chip_rows(chip, ev, words)
        assert rows is not None, "oracle/_ref/libzkref_core.so is missing: make -C oracle ref"
        out[chip] = {"width": int(rows.shape[1]), "events": ev.tolist(), "rows": rows.tolist()}
        print(chip, ev.shape, rows.shape)
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "more_rows.json"), "w"))

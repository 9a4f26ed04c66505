#!/usr/bin/env python3
"""How should column pieces of a pinned row-major trace cross PCIe?  (tools; not on the product path)"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ziren_b200 import _ffi, synthetic  # noqa: E402
from ziren_b200.prover import B200Prover  # noqa: E402

prover = B200Prover(synthetic.mini_case().machine)
lib = _ffi.lib()
rows, row_bytes = 1 << 17, int(sys.argv[1]) if len(sys.argv) > 1 else 16384           # 2 GiB, rows of 4096 words
total = rows * row_bytes


def run(mode, seg, n):
    ms = C.c_float()
    best = 1e9
    for _ in range(3):
        rc = lib.zkb200_h2d_probe(prover._h, mode, row_bytes, rows, seg, n, C.byref(ms))
        assert rc == 0, lib.zkb200_last_error(None)
        best = min(best, ms.value)
    return {"mode": ["dma2d", "pull", "dma"][mode], "seg_bytes": seg, "n": n, "ms": round(best, 2), "GB/s": round(total / best / 1e6, 1)}


for n in (1, 2, 4):
    print(json.dumps(run(2, 0, n)), flush=True)
for seg in (1024, 4096, row_bytes):
    for n in (1, 2):
        print(json.dumps(run(0, seg, n)), flush=True)
for n in (16, 64):
    print(json.dumps(run(1, 0, n)), flush=True)

#!/usr/bin/env python3
"""Which split of a two-level coset LDE is fastest?  For every log n the size K2 of the contiguous level
is forced through zkb200_set_option("ntt_k2") and the LDE timed in one process (about 2^27 elements)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ziren_b200 import _ffi, field as kb, synthetic  # noqa: E402
from ziren_b200.prover import B200Prover  # noqa: E402

prover = B200Prover(synthetic.mini_case().machine)
stream = torch.cuda.ExternalStream(prover.stream_ptr())


def timed(fn, reps=4):
    for _ in range(2):
        fn()
    prover.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


for log_n in range(13, 23):
    n = 1 << log_n
    w = max(1, (1 << 27) >> log_n)
    d_in = torch.randint(0, kb.P, (w, n), dtype=torch.int32, device="cuda")
    d_out = torch.empty((w, 2 * n), dtype=torch.int32, device="cuda")
    row = {"log_n": log_n, "width": w}
    for k2 in range(max(7, log_n - 12, (log_n + 1) // 2 - 2), min(12, log_n - 1) + 1):
        _ffi.lib().zkb200_set_option(b"ntt_k2", k2)
        row[f"K2={k2}"] = round(timed(lambda: prover.coset_lde(d_in, d_out, log_n, w, 1, 3)), 3)
    _ffi.lib().zkb200_set_option(b"ntt_k2", 0)
    print(json.dumps(row), flush=True)
    del d_in, d_out

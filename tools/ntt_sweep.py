#!/usr/bin/env python3
"""BASELINE.json configs[3]: NTT microbench, KoalaBear columns of 2^16 ... 2^24 rows, HBM GB/s against the
measured roofline, one GPU.  One process, one JSON line per (kind, log_n, width):
  python tools/ntt_sweep.py [--min 16] [--max 24] [--elems-log 27]
`width` is chosen so that every case transforms about 2^elems_log elements (well above the 126 MB L2)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ziren_b200 import field as kb, synthetic  # noqa: E402
from ziren_b200.prover import B200Prover  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--min", type=int, default=16)
ap.add_argument("--max", type=int, default=24)
ap.add_argument("--elems-log", type=int, default=27)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()

peak = 6549.4
pk_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk_path):
    peak = float(json.load(open(pk_path))["hbm_gbs"])
prover = B200Prover(synthetic.mini_case().machine)
stream = torch.cuda.ExternalStream(prover.stream_ptr())


def timed(fn):
    for _ in range(2):
        fn()
    prover.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.reps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / args.reps


for log_n in range(args.min, args.max + 1):
    n = 1 << log_n
    # BASELINE configs[3]: 1, 64 and 256 columns, plus the width that makes about 2^elems_log elements
    for w in sorted({1, 64, 256, max(1, (1 << args.elems_log) >> log_n)}):
        if n * w * 4 * 3 > (60 << 30):
            continue
        d_in = torch.randint(0, kb.P, (w, n), dtype=torch.int32, device="cuda")
        d_out = torch.empty((w, 2 * n), dtype=torch.int32, device="cuda")
        cases = [("ntt", lambda: prover.ntt(d_in, d_out, log_n, w, False, True), 8.0 * n * w)]
        if log_n + 1 <= 24:          # the LDE domain must fit the field's two-adicity (2^24)
            cases.append(("coset_lde_x2", lambda: prover.coset_lde(d_in, d_out, log_n, w, 1, 3), 12.0 * n * w))
        for kind, fn, alg in cases:
            ms = timed(fn)
            gbs = alg / (ms / 1e3) / 1e9
            print(json.dumps({"kind": kind, "log_n": log_n, "width": w, "ms": round(ms, 4), "algorithmic_GB": alg / 1e9, "GB/s": round(gbs, 1),
                              "frac_of_measured_hbm": round(gbs / peak, 4), "frac_of_8TBs": round(gbs / 8000.0, 4)}), flush=True)
        del d_in, d_out

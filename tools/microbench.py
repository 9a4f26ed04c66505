#!/usr/bin/env python3
"""Kernel-level micro-benchmarks through the C ABI (device-resident, column-major inputs), timed
with CUDA events on the prover's stream.  Used for the ncu captures under profiles/.
  python tools/microbench.py lde --log-n 18 --width 512 --reps 5
  python tools/microbench.py mmcs --log-n 19 --width 4167 --reps 3
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ziren_b200 import field as kb  # noqa: E402
from ziren_b200 import synthetic  # noqa: E402
from ziren_b200.prover import B200Prover  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("what", choices=["lde", "ntt", "mmcs", "permute", "tracegen", "keccak"])
ap.add_argument("--chip", default="ShiftRight", help="tracegen: any chip with a row filler (csrc/tracegen.cu alu_chip_by_name, or Global)")
ap.add_argument("--col-major", action="store_true", help="tracegen: write the column-major layout")
ap.add_argument("--log-n", type=int, default=18)
ap.add_argument("--width", type=int, default=512)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=2)
args = ap.parse_args()

prover = B200Prover(synthetic.mini_case().machine)
stream = torch.cuda.ExternalStream(prover.stream_ptr())
n, w = 1 << args.log_n, args.width
PEAK = 6542.1


def timed(fn):
    for _ in range(args.warmup):
        fn()
    prover.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.reps):
        fn()
    e1.record(stream)
    e1.synchronize()
    return e0.elapsed_time(e1) / args.reps


if args.what == "lde":
    d_in = torch.randint(0, kb.P, (w, n), dtype=torch.int32, device="cuda")
    d_out = torch.empty((w, 2 * n), dtype=torch.int32, device="cuda")
    ms = timed(lambda: prover.coset_lde(d_in, d_out, args.log_n, w, 1, 3))
    alg = 12.0 * n * w
elif args.what == "ntt":
    d_in = torch.randint(0, kb.P, (w, n), dtype=torch.int32, device="cuda")
    d_out = torch.empty_like(d_in)
    ms = timed(lambda: prover.ntt(d_in, d_out, args.log_n, w, False, True))
    alg = 8.0 * n * w
elif args.what == "mmcs":
    d_in = torch.randint(0, kb.P, (w, n), dtype=torch.int32, device="cuda")
    ms = timed(lambda: prover.mmcs_root([d_in], [args.log_n], [w]))
    alg = 4.0 * n * w + 32.0 * n + 96.0 * (n - 1)
elif args.what == "tracegen":
    from ziren_b200 import tracegen as tg
    w = tg.width(args.chip)
    gens = {"Mul": tg.synthetic_mul_events, "MemoryInstrs": tg.synthetic_mem_instr_events, "MemoryLocal": tg.synthetic_memory_local_events,
            "Cpu": tg.synthetic_cpu_events, "MiscInstrs": tg.synthetic_misc_events, "DivRem": tg.synthetic_div_rem_events,
            "Global": tg.synthetic_global_events,
            "SyscallInstrs": lambda k, seed: tg.synthetic_syscall_events(k, seed=seed, kind="instrs"),
            "SyscallCore": lambda k, seed: tg.synthetic_syscall_events(k, seed=seed, kind="core"),
            "MemoryGlobalInit": lambda k, seed: tg.memory_global_records(tg.synthetic_memory_global_events(k, seed=seed), 0)}
    n_ev = (n - 77) * tg.events_per_row(args.chip)
    ev = gens[args.chip](n_ev, seed=1) if args.chip in gens else tg.synthetic_events(args.chip, n_ev, seed=1)
    d_ev = torch.from_numpy(ev.view(np.int32)).cuda()
    d_out = torch.empty((n * w,), dtype=torch.int32, device="cuda")
    ms = timed(lambda: prover.generate_alu_trace(args.chip, d_ev, args.log_n, d_out, col_major=args.col_major))
    alg = 4.0 * ev.size + 4.0 * n * w            # the event records read, the rows written (Global is compute-bound: see Grows/s)
elif args.what == "keccak":
    from ziren_b200 import keccak_sponge as ksp
    w = ksp.WIDTH
    nb = n // 24 - 3
    blocks = ksp.synthetic_blocks(nb // 4, 4, seed=1)
    d_ev = torch.from_numpy(blocks.view(np.int32)).cuda()
    d_out = torch.empty((n * w,), dtype=torch.int32, device="cuda")
    ms = timed(lambda: prover.generate_keccak_sponge_trace(d_ev, args.log_n, d_out, col_major=True))
    alg = 1536.0 * len(blocks) + 4.0 * n * w     # one record read per block, one row written
else:
    d_in = torch.randint(0, kb.P, (n, 16), dtype=torch.int32, device="cuda")
    ms = timed(lambda: (prover.poseidon2_permute_batch(d_in, n), prover.sync()))
    alg = 128.0 * n
gbs = alg / (ms / 1e3) / 1e9
out = {"what": args.what, "log_n": args.log_n, "width": w, "ms": ms, "algorithmic_GB": alg / 1e9, "GB/s": gbs, "frac_of_measured_hbm": gbs / PEAK}
if args.what == "keccak":
    out.update({"chip": "KeccakSponge", "layout": "column-major", "Grows/s": n / (ms / 1e3) / 1e9, "PCIe_ms_for_the_same_rows_at_55GB/s": 4.0 * n * w / 55e9 * 1e3})
if args.what == "tracegen":
    out.update({"chip": args.chip, "width": w, "layout": "column-major" if args.col_major else "row-major", "Grows/s": n / (ms / 1e3) / 1e9})
if args.what in ("mmcs", "permute"):
    perms = n * (-(-w // 8)) + (n - 1) if args.what == "mmcs" else n
    out["Gperm/s"] = perms / (ms / 1e3) / 1e9
print(json.dumps(out))

#!/usr/bin/env python3
"""K3 alone on the real KeccakSponge chip (ziren_b200/keccak_air.py): the generated constraint kernel `qk` on random
LDEs of a 2^log_n-row table (the kernel does the same work whether or not the constraints hold), timed with CUDA
events on the prover's stream; the command ncu wraps for profiles/*qk*.  usage: qk_probe.py [log_n] [reps]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ziren_b200 import field as kb  # noqa: E402
from ziren_b200 import keccak_sponge as ks  # noqa: E402
from ziren_b200 import synthetic  # noqa: E402
from ziren_b200.prover import B200Prover  # noqa: E402

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if len(sys.argv) > 3:       # threads per CTA of the generated kernel
    from ziren_b200 import _ffi
    _ffi.lib().zkb200_set_option(b"qk_block", int(sys.argv[3]))
case = synthetic.keccak_real_case(ks.synthetic_blocks(1, 1), None, log_cpu=10)
chip = case.machine.chip("KeccakSponge")
prover = B200Prover(case.machine)
stream = torch.cuda.ExternalStream(prover.stream_ptr())
H = 2 << log_n
main = torch.randint(0, kb.P, (chip.main_width, H), dtype=torch.int32, device="cuda")
perm = torch.randint(0, kb.P, (4 * chip.perm_width_ef, H), dtype=torch.int32, device="cuda")
out = torch.empty((2, 4, 1 << log_n), dtype=torch.int32, device="cuda")
rng = np.random.default_rng(1)
ef = lambda: rng.integers(0, kb.P, 4, dtype=np.uint32)
args = (ef(), ef(), ef(), rng.integers(0, kb.P, 14, dtype=np.uint32), ef(), case.public_values)


def run():
    prover.quotient("KeccakSponge", log_n, None, main, perm, *args, out)


run()
prover.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(reps):
    run()
e1.record(stream)
e1.synchronize()
ms = e0.elapsed_time(e1) / reps
alg = 4.0 * H * (chip.main_width + 4 * chip.perm_width_ef) + 16.0 * H
print(json.dumps({"what": "qk KeccakSponge", "log_n": log_n, "ms": ms, "algorithmic_GB": alg / 1e9, "GB/s": alg / ms / 1e6,
                  "constraints": chip.num_constraints, "lookups": chip.num_local_lookups}))

#!/usr/bin/env python3
"""Setup + N shard proofs of the keccak workload, nothing else: the command ncu wraps for the
per-launch time list under profiles/.  usage: one_step.py [log_cpu] [n_proofs] [real]
`real`: the bench's default workload (the REAL KeccakSponge chip, its table generated inside the commit from event
records); otherwise the synthetic 4167-column stand-in."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ziren_b200 import field as kb, synthetic
from ziren_b200.prover import B200Prover

log_cpu = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
real = len(sys.argv) > 3 and sys.argv[3] == "real"
if real:
    from ziren_b200 import keccak_sponge as ksp
    from ziren_b200.prover import EventTrace
    rows = 1 << (log_cpu - 2)
    blocks = ksp.synthetic_blocks(max(1, rows // 96), 4 if rows >= 96 else 1)
    case = synthetic.keccak_real_case(blocks, None, log_cpu=log_cpu)
else:
    case = synthetic.keccak_case(log_cpu=log_cpu)
prover = B200Prover(case.machine)
pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
dev = {k: torch.from_numpy(kb.to_monty(v).view(np.int32)).cuda() for k, v in case.traces.items()}
if real:
    ev = torch.from_numpy(blocks.view(np.int32)).cuda()
    dev["KeccakSponge"] = EventTrace(ev, log_cpu - 2, ksp.WIDTH)
for i in range(n):
    l0 = prover.launch_count()
    proof, _ = prover.prove_shard(pk, dev, case.public_values)
    print("proof", i, proof.size, "words;", prover.launch_count() - l0, "kernel launches")

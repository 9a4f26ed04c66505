#!/usr/bin/env python3
"""A/B of run-time options on the per-stage device times of one bench-size proof (one process, data made once).
usage: stage_ab.py [log_cpu] key=value[,key=value] ...   e.g. stage_ab.py 20 k4b_rows=2 k4b_rows=4 qk_block=256"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ziren_b200 import _ffi, field as kb, synthetic  # noqa: E402
from ziren_b200.prover import B200Prover  # noqa: E402

log_cpu = int(sys.argv[1]) if len(sys.argv) > 1 else 20
variants = [""] + sys.argv[2:]
case = synthetic.keccak_case(log_cpu=log_cpu)
prover = B200Prover(case.machine)
pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
dev = {k: torch.from_numpy(kb.to_monty(v).view(np.int32)).cuda() for k, v in case.traces.items()}
defaults = {"k4b_rows": 2, "qk_block": 256, "quotient_codegen": 1, "eval_v2": -1, "ntt_lean": 3}
for _ in range(2):
    prover.prove_shard(pk, dev, case.public_values)
for v in variants:
    for k, d in defaults.items():
        _ffi.lib().zkb200_set_option(k.encode(), d)
    for kv in filter(None, v.split(",")):
        k, val = kv.split("=")
        assert _ffi.lib().zkb200_set_option(k.encode(), int(val)) == 0, k
    prover.prove_shard(pk, dev, case.public_values)
    prover.set_profile(True)
    prover.prove_shard(pk, dev, case.public_values)
    st = prover.last_stage_times()
    prover.set_profile(False)
    print(json.dumps({"variant": v or "default", "total": round(sum(st.values()), 2), **{k: round(x, 2) for k, x in st.items()}}), flush=True)

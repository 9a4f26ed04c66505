#!/usr/bin/env python3
"""Per-call wall times of concurrent prove_shard calls with pinned host inputs (diagnostic)."""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ziren_b200 import field as kb, synthetic
from ziren_b200.prover import B200Prover

log_cpu = int(sys.argv[1]) if len(sys.argv) > 1 else 19
nthreads = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
case = synthetic.keccak_case(log_cpu=log_cpu)
prover = B200Prover(case.machine)
pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
base = pk.observe_into()
host = {k: torch.from_numpy(kb.to_monty(v).view(np.int32)).pin_memory() for k, v in case.traces.items()}
t00 = time.perf_counter()
log = []
def worker(t):
    for i in range(t, steps, nthreads):
        t0 = time.perf_counter()
        data = prover.commit(host, case.public_values)
        t1 = time.perf_counter()
        prover.open(pk, data, base)
        t2 = time.perf_counter()
        data.free()
        log.append((i, t, round(1e3 * (t0 - t00), 1), round(1e3 * (t1 - t0), 1), round(1e3 * (t2 - t1), 1)))
ths = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
[t.start() for t in ths]; [t.join() for t in ths]
tot = time.perf_counter() - t00
for r in sorted(log): print("step %d thread %d start %.1f commit %.1f open %.1f" % r)
print("total %.1f ms, %.1f ms/step" % (1e3 * tot, 1e3 * tot / steps))
free, total = torch.cuda.mem_get_info()
print("device memory in use: %.1f GB" % ((total - free) / 1e9))

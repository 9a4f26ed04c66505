"""Fast GPU check of the MemoryLocal row kernel without importing torch (see tools/gpu_check_mem_instr.py): the pytest cases
of tests/test_zz_tracegen_memory_local.py through the C ABI against the oracle and the reference-written golden entries."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_ffi as orc  # noqa: E402
from ziren_b200 import field as kb  # noqa: E402
from ziren_b200 import synthetic  # noqa: E402
from ziren_b200 import tracegen as tg  # noqa: E402
from ziren_b200.prover import B200Prover, ZkbError  # noqa: E402

rt = C.CDLL("libcudart.so.12")
rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]


class DevWords:
    def __init__(self, n):
        p = C.c_void_p()
        assert rt.cudaMalloc(C.byref(p), 4 * n) == 0
        self.p, self.shape = p.value, (n,)

    def data_ptr(self):
        return self.p

    def numpy(self):
        out = np.empty(self.shape[0], np.uint32)
        assert rt.cudaMemcpy(out.ctypes.data, self.p, 4 * self.shape[0], 2) == 0
        return out


prover = B200Prover(synthetic.mini_case().machine, device=0)
bad, w = 0, 56
for n, log_h, cm in ((20000, 13, False), (20001, 13, True), (513, 8, True), (1, 4, False), (0, 4, True), (64, 4, False)):
    ev, h = tg.synthetic_memory_local_events(n, seed=20 + n), 1 << log_h
    out = DevWords(h * w)
    prover.generate_alu_trace("MemoryLocal", ev, log_h, out, col_major=cm)
    got = out.numpy()
    got = got.reshape(w, h).T if cm else got.reshape(h, w)
    ok = np.array_equal(got, kb.to_monty(orc.memory_local_trace(ev, h)))
    bad += not ok
    print("MemoryLocal", n, log_h, cm, "ok" if ok else "MISMATCH", flush=True)
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "memory_local_entries.json")))
gev, gent = np.array(gold["events"], np.uint32), np.array(gold["entries"], np.uint32)
out = DevWords(32 * w)
prover.generate_alu_trace("MemoryLocal", gev, 5, out)
flat = np.zeros((128, 14), np.uint32)
flat[: len(gent)] = gent
ok = np.array_equal(out.numpy().reshape(32, w), flat.reshape(32, w))
bad += not ok
print("golden", "ok" if ok else "MISMATCH", flush=True)
try:
    prover.generate_alu_trace("MemoryLocal", tg.synthetic_memory_local_events(20000), 12, DevWords(4096 * w))
    bad += 1
    print("too many events: NOT refused")
except ZkbError as e:
    print("too many events refused:", e)
print("%.1fs" % (time.time() - t0), flush=True)
prover.close()
sys.exit(1 if bad else 0)

"""Fast GPU check of the MemoryInstrs and Mul row kernels without importing torch (device memory through libcudart
by ctypes): rows through the C ABI against the oracle and the reference-written golden rows.  Exit code 0 = bit-exact."""
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
from ziren_b200.prover import B200Prover  # noqa: E402

rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
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
print("ctx %.1fs" % (time.time() - t0), flush=True)
bad = 0
for chip, events, trace in (("MemoryInstrs", tg.synthetic_mem_instr_events, orc.mem_instr_trace), ("Mul", tg.synthetic_mul_events, orc.mul_trace)):
    w = tg.width(chip)
    for n, log_h, cm in ((5000, 13, False), (5000, 13, True), (129, 8, True), (1, 4, False), (0, 4, True)):
        ev, h = events(n, seed=20 + n), 1 << log_h
        out = DevWords(h * w)
        prover.generate_alu_trace(chip, ev, log_h, out, col_major=cm)
        got = out.numpy()
        got = got.reshape(w, h).T if cm else got.reshape(h, w)
        ok = np.array_equal(got, kb.to_monty(trace(ev, h)))
        bad += not ok
        print(chip, n, log_h, cm, "ok" if ok else "MISMATCH", flush=True)
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "mem_instr_rows.json")))
gev, grows = np.array(gold["events"], np.uint32), np.array(gold["rows"], np.uint32)
out = DevWords(128 * 79)
prover.generate_alu_trace("MemoryInstrs", gev, 7, out)
ok = np.array_equal(out.numpy().reshape(128, 79)[: len(gev)], grows)
bad += not ok
print("golden", "ok" if ok else "MISMATCH", "%.1fs" % (time.time() - t0), flush=True)
prover.close()
sys.exit(1 if bad else 0)

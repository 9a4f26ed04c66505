"""Fast GPU check of the event -> row kernels (K6) without importing torch (device memory through libcudart by ctypes, so a
run costs seconds of box time): for every chip with a row filler, the pytest cases of tests/test_tracegen*.py and
tests/test_zz_tracegen_*.py through the C ABI against the oracle, plus the reference-written golden rows.
Usage: python tools/gpu_check_tracegen.py [chip ...]   (default: all).  Exit code 0 = bit-exact."""
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


def golden(name, rows_key="rows"):
    g = json.load(open(os.path.join(ROOT, "tests", "golden", name)))
    return np.array(g["events"], np.uint32), np.array(g[rows_key], np.uint32)


CHIPS = {c: (lambda n, s, c=c: tg.synthetic_events(c, n, seed=s), lambda ev, h, c=c: orc.alu_trace(c, ev, h)) for c in tg.ALU_CHIPS}
CHIPS["Mul"] = (lambda n, s: tg.synthetic_mul_events(n, seed=s), orc.mul_trace)
CHIPS["MemoryInstrs"] = (lambda n, s: tg.synthetic_mem_instr_events(n, seed=s), orc.mem_instr_trace)
CHIPS["MemoryLocal"] = (lambda n, s: tg.synthetic_memory_local_events(n, seed=s), orc.memory_local_trace)
CHIPS["Cpu"] = (lambda n, s: tg.synthetic_cpu_events(n, seed=s), orc.cpu_trace)
CHIPS["MiscInstrs"] = (lambda n, s: tg.synthetic_misc_events(n, seed=s), orc.misc_trace)
CHIPS["DivRem"] = (lambda n, s: tg.synthetic_div_rem_events(n, seed=s), lambda ev, h: orc.chip_trace("DivRem", ev, h))
for _chip, _kind in (("SyscallCore", "core"), ("SyscallPrecompile", "precompile"), ("SyscallInstrs", "instrs")):
    CHIPS[_chip] = (lambda n, s, k=_kind: tg.synthetic_syscall_events(n, seed=s, kind=k), lambda ev, h, c=_chip: orc.chip_trace(c, ev, h))


def _memory_global_trace(records, h):
    """The oracle takes the sorted events and the previous address; both are in the flattened records."""
    return orc.memory_global_trace(records[:, :4], int(records[0, 4]) if len(records) else 0, h)


for _chip in ("MemoryGlobalInit", "MemoryGlobalFinalize"):
    CHIPS[_chip] = (lambda n, s: tg.memory_global_records(tg.synthetic_memory_global_events(n, seed=s), 0 if s % 2 else 5), _memory_global_trace)
CHIPS["Global"] = (lambda n, s: tg.synthetic_global_events(n, seed=s), orc.global_trace)
GOLDEN = {"Mul": "mul_rows.json", "MemoryInstrs": "mem_instr_rows.json", "Cpu": "cpu_rows.json", "MiscInstrs": "misc_rows.json"}

prover = B200Prover(synthetic.mini_case().machine, device=0)
bad = 0
for chip in (sys.argv[1:] or list(CHIPS)):
    events, trace = CHIPS[chip]
    w, epr = tg.width(chip), tg.events_per_row(chip)
    for n, log_h, cm in ((5000 * epr, 13, False), (5000 * epr + 1, 13, True), (129, 8, True), (1, 4, False), (0, 4, True)):
        ev, h = events(n, 20 + n), 1 << log_h
        out = DevWords(h * w)
        prover.generate_alu_trace(chip, ev, log_h, out, col_major=cm)
        got = out.numpy()
        got = got.reshape(w, h).T if cm else got.reshape(h, w)
        ok = np.array_equal(got, kb.to_monty(trace(ev, h)))
        bad += not ok
        print(chip, n, log_h, cm, "ok" if ok else "MISMATCH", flush=True)
    if chip in GOLDEN:
        gev, grows = golden(GOLDEN[chip])
        out = DevWords(128 * w)
        prover.generate_alu_trace(chip, gev, 7, out)
        ok = np.array_equal(out.numpy().reshape(128, w)[: len(gev)], grows)
        bad += not ok
        print(chip, "golden", "ok" if ok else "MISMATCH", flush=True)
    if chip == "MemoryLocal":
        gev, gent = golden("memory_local_entries.json", "entries")
        out = DevWords(32 * w)
        prover.generate_alu_trace(chip, gev, 5, out)
        flat = np.zeros((128, 14), np.uint32)
        flat[: len(gent)] = gent
        ok = np.array_equal(out.numpy().reshape(32, w), flat.reshape(32, w))
        bad += not ok
        print(chip, "golden", "ok" if ok else "MISMATCH", flush=True)
    try:
        prover.generate_alu_trace(chip, events(16 * epr + 1, 1), 4, DevWords(16 * w))
        bad += 1
        print(chip, "too many events: NOT refused")
    except ZkbError:
        pass
print("%.1fs, %d mismatches" % (time.time() - t0, bad), flush=True)
prover.close()
sys.exit(1 if bad else 0)

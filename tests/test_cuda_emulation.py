"""The launchers and __global__ wrappers of the row fillers (K6 alu_rows_kernel for twenty chips, K6b KeccakSponge, K6c Global:
lift, recursive curve-point scan, finish) and of K7 (derive_multiplicities), run on the CPU from the product's own source text.

tests/cudaemu/build.py compiles csrc/tracegen.cu, csrc/derive.cu and csrc/machine.cpp for the host against a stand-in of the
CUDA runtime and execution model (tests/cudaemu/cuda_runtime.h: blocks one after the other, the threads of a block as real
threads that meet at __syncthreads, `__shared__` tiles, `__constant__` tables); only the `<<<...>>>` syntax is rewritten.
What tests/hostcheck walks are the device FUNCTIONS; this covers what is around them - grid sizes, the shared-memory tile and
its phases, bounds checks of partial CTAs, scratch buffers, the scan's recursion, the host logic that picks receives and sends.
Nine of these chips and K7 were written after the round's GPU budget was spent: this and hostcheck are what stands behind
them until their GPU cases (sorted last in the GPU run) have run.  Every comparison is bit-exact against the oracle.
"""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from ziren_b200 import field as kb
from ziren_b200 import keccak_sponge as ks
from ziren_b200 import synthetic
from ziren_b200 import tracegen as tg

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    spec = importlib.util.spec_from_file_location("cudaemu_build", os.path.join(HERE, "cudaemu", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = C.CDLL(mod.build())
    lib.emu_last_error.restype = C.c_char_p
    lib.emu_launches.restype = C.c_ulonglong
    return lib


def _chips(orc):
    chips = {c: (lambda n, s, c=c: tg.synthetic_events(c, n, seed=s), lambda ev, h, c=c: orc.alu_trace(c, ev, h)) for c in tg.ALU_CHIPS}
    chips["Mul"] = (lambda n, s: tg.synthetic_mul_events(n, seed=s), orc.mul_trace)
    chips["MemoryInstrs"] = (lambda n, s: tg.synthetic_mem_instr_events(n, seed=s), orc.mem_instr_trace)
    chips["MemoryLocal"] = (lambda n, s: tg.synthetic_memory_local_events(n, seed=s), orc.memory_local_trace)
    chips["Cpu"] = (lambda n, s: tg.synthetic_cpu_events(n, seed=s), orc.cpu_trace)
    chips["MiscInstrs"] = (lambda n, s: tg.synthetic_misc_events(n, seed=s), orc.misc_trace)
    chips["DivRem"] = (lambda n, s: tg.synthetic_div_rem_events(n, seed=s), lambda ev, h: orc.chip_trace("DivRem", ev, h))
    for chip, kind in (("SyscallCore", "core"), ("SyscallPrecompile", "precompile"), ("SyscallInstrs", "instrs")):
        chips[chip] = (lambda n, s, k=kind: tg.synthetic_syscall_events(n, seed=s, kind=k), lambda ev, h, c=chip: orc.chip_trace(c, ev, h))
    for chip in ("MemoryGlobalInit", "MemoryGlobalFinalize"):
        chips[chip] = (lambda n, s: tg.memory_global_records(tg.synthetic_memory_global_events(n, seed=s), 0 if s % 2 else 5),
                       lambda rec, h: orc.memory_global_trace(rec[:, :4], int(rec[0, 4]) if len(rec) else 0, h))
    chips["Global"] = (lambda n, s: tg.synthetic_global_events(n, seed=s), orc.global_trace)
    return chips


ALL_CHIPS = tuple(tg.ALU_CHIPS) + ("Mul", "MemoryInstrs", "MemoryLocal", "Cpu", "MiscInstrs", "DivRem", "SyscallCore", "SyscallPrecompile",
                                   "SyscallInstrs", "MemoryGlobalInit", "MemoryGlobalFinalize", "Global")


def _generate(emu, chip, ev, h, col_major):
    w = emu.emu_chip_width(chip.encode())
    assert w == (ks.WIDTH if chip == "KeccakSponge" else tg.width(chip))
    ev = np.ascontiguousarray(ev)
    out = np.full(h * w, 0xFFFFFFFF, np.uint32)
    rc = emu.emu_generate_trace(chip.encode(), C.c_void_p(ev.ctypes.data), C.c_size_t(len(ev)), C.c_size_t(h), C.c_void_p(out.ctypes.data),
                                C.c_int(int(col_major)))
    if rc:
        raise RuntimeError(emu.emu_last_error().decode())
    return out.reshape(w, h).T if col_major else out.reshape(h, w)


@pytest.mark.parametrize("chip", ALL_CHIPS)
def test_row_filler_launchers_match_the_oracle(emu, oracle, chip):
    """Full CTAs, a partial last CTA, a table shorter than one CTA, one event, no event; both output layouts."""
    events, trace = _chips(oracle)[chip]
    epr = tg.events_per_row(chip)
    before = emu.emu_launches()
    for n, log_h, cm in ((300 * epr, 9, False), (300 * epr + 1, 9, True), (129, 8, True), (1, 4, False), (0, 4, True), (3, 1, True)):
        if n > (1 << log_h) * epr:
            n = (1 << log_h) * epr
        ev, h = events(n, 20 + n), 1 << log_h
        got = _generate(emu, chip, ev, h, cm)
        assert np.array_equal(got, kb.to_monty(trace(ev, h))), (chip, n, log_h, cm)
    assert emu.emu_launches() > before
    with pytest.raises(RuntimeError, match="more events"):
        _generate(emu, chip, events(16 * epr + 1, 1), 16, True)


def test_global_scan_recursion_levels(emu, oracle):
    """The chunked scan with one, two and three levels (chunks of 32 points: 33, 1025 and 1100 points with the start point)."""
    for n, log_h in ((31, 5), (32, 6), (1024, 10), (1099, 11)):
        ev = tg.synthetic_global_events(n, seed=n)
        got = _generate(emu, "Global", ev, 1 << log_h, True)
        assert np.array_equal(got, kb.to_monty(oracle.global_trace(ev, 1 << log_h))), n


def test_keccak_sponge_launcher_matches_the_oracle(emu, oracle):
    for per in ([1], [2, 1, 3]):
        b = ks.synthetic_blocks(len(per), per, seed=3, shard=1)
        for extra in (0, 1):
            h = 1 << (ks.padded_log_height(len(b)) + extra)
            got = _generate(emu, "KeccakSponge", b.reshape(len(b), -1), h, True)
            assert np.array_equal(got, kb.to_monty(oracle.keccak_sponge_trace(b, h)))
    with pytest.raises(RuntimeError, match="more rows"):
        _generate(emu, "KeccakSponge", ks.synthetic_blocks(2, [1, 1], seed=1).reshape(2, -1), 32, True)


# ---- K7 through csrc/derive.cu's own host logic and kernels ---------------------------------------------------------------
def _derive(emu, case, receiver):
    desc = np.ascontiguousarray(case.machine.descriptor(), dtype=np.uint32)
    cm = lambda a: np.ascontiguousarray(kb.to_monty(np.ascontiguousarray(np.asarray(a).T)))
    names = [n for n in case.traces if n != receiver]
    preps = [cm(case.prep[n]) if n in case.prep else None for n in names]
    mains = [cm(case.traces[n]) for n in names]
    k = len(names)
    na = (C.c_char_p * k)(*[n.encode() for n in names])
    pp = (C.c_void_p * k)(*[(p.ctypes.data if p is not None else None) for p in preps])
    mp = (C.c_void_p * k)(*[m.ctypes.data for m in mains])
    hs = (C.c_size_t * k)(*[case.traces[n].shape[0] for n in names])
    rprep = cm(case.prep[receiver])
    h, w = case.prep[receiver].shape[0], case.machine.chip(receiver).main_width
    out = np.full((w, h), 0xFFFFFFFF, np.uint32)
    n_lookups = C.c_ulonglong(0)
    rc = emu.emu_derive(C.c_void_p(desc.ctypes.data), C.c_size_t(desc.size), receiver.encode(), C.c_void_p(rprep.ctypes.data), C.c_size_t(h),
                        C.c_int(k), na, pp, mp, hs, C.c_void_p(out.ctypes.data), C.byref(n_lookups))
    if rc:
        raise RuntimeError(emu.emu_last_error().decode())
    return kb.from_monty(out).T, int(n_lookups.value)


def _senders(case, receiver):
    return [(n, case.prep.get(n), t) for n, t in case.traces.items() if n != receiver]


@pytest.mark.parametrize("which", ["mini", "edge", "core"])
def test_derive_multiplicities_launcher_matches_the_oracle(emu, oracle, which):
    case = {"mini": synthetic.mini_case, "edge": synthetic.edge_case,
            "core": lambda: synthetic.fibonacci_core_case(log_cpu=8, num_queries=4, pow_bits=2)}[which]()
    om = oracle.OracleMachine(case.machine)
    for receiver in ("Byte", "Program"):
        if receiver not in case.traces:
            continue
        want, n = om.derive_multiplicities(receiver, case.prep[receiver], _senders(case, receiver), case.traces[receiver].shape[1])
        got, counted = _derive(emu, case, receiver)
        assert counted == n
        assert np.array_equal(got, want), receiver
        assert np.array_equal(got, case.traces[receiver]), receiver


def test_derive_multiplicities_launcher_on_the_real_keccak_shard(emu, oracle):
    b = ks.synthetic_blocks(2, [1, 1], seed=5, shard=1)
    t = oracle.keccak_sponge_trace(b, 1 << ks.padded_log_height(len(b)))
    case = synthetic.keccak_real_case(b, t, log_cpu=8, num_queries=4, pow_bits=2)
    got, _ = _derive(emu, case, "Byte")
    assert np.array_equal(got, case.traces["Byte"])


def test_derive_multiplicities_launcher_refuses(emu):
    case = synthetic.mini_case()
    bad = {k: v.copy() for k, v in case.traces.items()}
    name = next(n for n in bad if n not in ("Byte", "Program"))
    bad[name][1, 5] = 1 << synthetic.RANGE_BITS
    with pytest.raises(RuntimeError, match="1 lookups are in no row of Byte"):
        _derive(emu, synthetic.ShardCase(case.machine, case.prep, bad, case.public_values, case.cycles), "Byte")
    # a chip whose receives are not made of preprocessed columns cannot be derived
    with pytest.raises(RuntimeError, match="has no receive of preprocessed columns"):
        c2 = synthetic.ShardCase(case.machine, {**case.prep, "Cpu": np.zeros((64, 0), np.uint32)}, case.traces, case.public_values, case.cycles)
        _derive(emu, c2, "Cpu")

"""Trace generation of DivRem, SyscallCore, SyscallPrecompile, SyscallInstrs, MemoryGlobalInit and MemoryGlobalFinalize
(SURVEY.md section 8 row f3).

CPU tests: the oracle (oracle/tracegen.h div_rem_row, syscall_row, syscall_instr_row, memory_global_trace) and the product's row
fillers compiled for the host (ziren_b200/csrc/tracegen.cuh fill_div_rem, fill_syscall, fill_syscall_instr,
fill_memory_global), row by row and through the CTA phases the CUDA kernel runs (AluCta: load / fill / store with the
kernel's own index arithmetic), against golden rows written by the REFERENCE'S OWN C++ (crates/core/machine/include/{div_rem,
syscall,syscall_instrs,memory_global}.hpp; tests/golden/more_rows.json) and, when oracle/_ref is present, against that C++
live; the rows against the executor's semantics.  GPU: the CUDA kernel through the C ABI."""
import ctypes
import json
import os

import numpy as np
import pytest

from ziren_b200 import field as kb
from ziren_b200 import tracegen as tg

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "more_rows.json")))
# AluChip ids, csrc/tracegen.cuh
CHIP_ID = {"DivRem": 14, "SyscallCore": 15, "SyscallPrecompile": 16, "SyscallInstrs": 17, "MemoryGlobalInit": 18, "MemoryGlobalFinalize": 19}
ROW_CHIPS = ("DivRem", "SyscallCore", "SyscallPrecompile", "SyscallInstrs")
# the columns memory_global.hpp fills (shard, timestamp, addr, addr_bits, value, is_real); the rest comes from the neighbouring
# event in the reference's sequential loop (memory/global.rs:150-180)
MG_TWIN_COLS = np.r_[0:3, 35:106]


def synthetic(chip, n, seed):
    if chip == "DivRem":
        return tg.synthetic_div_rem_events(n, seed=seed)
    return tg.synthetic_syscall_events(n, seed=seed, kind={"SyscallCore": "core", "SyscallPrecompile": "precompile", "SyscallInstrs": "instrs"}[chip])


def twin_domain(chip, ev):
    """Events on which the reference's C++ twin and its Rust generate_trace agree (div_rem.hpp: c != 0, not INT_MIN / -1)."""
    if chip != "DivRem":
        return ev
    b, c = ev[:, 7], ev[:, 8]
    return ev[(c != 0) & ~((b == 0x80000000) & (c == 0xFFFFFFFF))]


def _host_rows(host, chip, ev, height, cta=False, col_major=False):
    w = tg.width(chip)
    ev = np.ascontiguousarray(ev, dtype=np.uint32).reshape(-1, tg.event_words(chip))
    out = np.full(height * w, 0xFFFFFFFF, np.uint32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    if cta:
        rc = host.hostcheck_alu_rows_cta(CHIP_ID[chip], p(ev), ctypes.c_size_t(len(ev)), ctypes.c_size_t(height), p(out), int(col_major))
    else:
        rc = host.hostcheck_alu_rows(CHIP_ID[chip], p(ev), ctypes.c_size_t(len(ev)), ctypes.c_size_t(height), p(out))
    assert rc == 0
    return out.reshape(w, height).T if col_major else out.reshape(height, w)


def test_shapes(oracle, host):
    assert host.hostcheck_alu_nchips() == 20
    for chip, (w, ew) in tg.ROW_CHIPS.items():
        assert tg.width(chip) == host.hostcheck_alu_width(CHIP_ID[chip]) == w
        assert tg.event_words(chip) == host.hostcheck_alu_event_words(CHIP_ID[chip]) == ew
    for chip in ROW_CHIPS:
        assert oracle.chip_trace_width(chip) == tg.width(chip) == GOLD[chip]["width"]
        assert oracle.chip_event_words(chip) == tg.event_words(chip)
    assert GOLD["MemoryGlobalInit"]["width"] == oracle.MEMGLOBAL_WIDTH == 111
    # the C ABI knows every chip by its MachineAir::name (no GPU needed for this call)
    from ziren_b200 import _ffi
    for chip in list(tg.ROW_CHIPS) + ["Global", "Cpu", "MiscInstrs", "MemoryLocal", "Mul", "AddSub"]:
        assert _ffi.lib().zkb200_alu_trace_width(chip.encode()) == tg.width(chip), chip
    assert _ffi.lib().zkb200_alu_trace_width(b"Byte") == -1
    # the wide chips run 64 rows per CTA so that events + row tile fit 48 KB of static shared memory, the others 128
    assert host.hostcheck_alu_cta_rows(CHIP_ID["DivRem"]) == host.hostcheck_alu_cta_rows(CHIP_ID["MemoryGlobalInit"]) == 64
    assert host.hostcheck_alu_cta_rows(CHIP_ID["SyscallInstrs"]) == host.hostcheck_alu_cta_rows(12) == 128


@pytest.mark.parametrize("chip", ROW_CHIPS)
def test_oracle_and_product_match_reference_golden_rows(oracle, host, chip):
    ev, rows = np.array(GOLD[chip]["events"], np.uint32), np.array(GOLD[chip]["rows"], np.uint32)
    assert np.array_equal(kb.to_monty(oracle.chip_trace(chip, ev, 128))[: len(ev)], rows)
    assert np.array_equal(_host_rows(host, chip, ev, 128)[: len(ev)], rows)


@pytest.mark.parametrize("chip", ROW_CHIPS)
def test_oracle_and_product_match_reference_cpp_live(oracle, host, chip):
    ev = twin_domain(chip, synthetic(chip, 9000, 3))
    ref = oracle.ref_chip_rows(chip, ev, tg.event_words(chip))
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference checkout on this machine)")
    orc = kb.to_monty(oracle.chip_trace(chip, ev, 16384))
    assert np.array_equal(orc[: len(ev)], ref)
    assert np.array_equal(_host_rows(host, chip, ev, 16384), orc)         # zero padding rows included


@pytest.mark.parametrize("chip", ROW_CHIPS)
@pytest.mark.parametrize("n,height", [(0, 16), (1, 16), (127, 128), (128, 128), (129, 256), (1000, 1024), (65, 64 * 3)])
def test_product_matches_oracle_through_the_cta_phases(oracle, host, chip, n, height):
    """Every event, the rows the C++ twin cannot pin included (DivRem by zero, INT_MIN / -1), through the kernel's CTA phases."""
    ev = synthetic(chip, n, 40 + n)
    want = kb.to_monty(oracle.chip_trace(chip, ev, height))
    assert np.array_equal(_host_rows(host, chip, ev, height), want)
    assert np.array_equal(_host_rows(host, chip, ev, height, cta=True), want)
    assert np.array_equal(_host_rows(host, chip, ev, height, cta=True, col_major=True), want)


def test_cta_phases_agree_with_the_plain_row_loop_for_every_chip(host):
    """All twenty chips of alu_rows_kernel: the CTA walk (load / fill / store as the kernel indexes them, 128 or 64 rows per
    CTA, both layouts, a ragged last CTA and a table shorter than one CTA) against the row-by-row loop over the same fillers."""
    names = ["AddSub", "Bitwise", "Lt", "ShiftLeft", "ShiftRight", "CloClz", "Branch", "Jump", "MovCond", "Mul", "MemoryInstrs", "MemoryLocal",
             "Cpu", "MiscInstrs", "DivRem", "SyscallCore", "SyscallPrecompile", "SyscallInstrs", "MemoryGlobalInit", "MemoryGlobalFinalize"]
    gens = {"Mul": tg.synthetic_mul_events, "MemoryInstrs": tg.synthetic_mem_instr_events, "MemoryLocal": tg.synthetic_memory_local_events,
            "Cpu": tg.synthetic_cpu_events, "MiscInstrs": tg.synthetic_misc_events, "DivRem": tg.synthetic_div_rem_events,
            "SyscallCore": lambda n, seed: tg.synthetic_syscall_events(n, seed=seed, kind="core"),
            "SyscallPrecompile": lambda n, seed: tg.synthetic_syscall_events(n, seed=seed, kind="precompile"),
            "SyscallInstrs": lambda n, seed: tg.synthetic_syscall_events(n, seed=seed, kind="instrs"),
            "MemoryGlobalInit": lambda n, seed: tg.memory_global_records(tg.synthetic_memory_global_events(n, seed=seed), 0),
            "MemoryGlobalFinalize": lambda n, seed: tg.memory_global_records(tg.synthetic_memory_global_events(n, seed=seed), 0)}
    assert host.hostcheck_alu_nchips() == len(names)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for cid, chip in enumerate(names):
        w, epr = tg.width(chip), tg.events_per_row(chip)
        assert host.hostcheck_alu_width(cid) == w and host.hostcheck_alu_event_words(cid) == tg.event_words(chip), chip
        for rows, height in ((300, 384), (40, 48)):
            n = rows * epr - (1 if epr > 1 else 0)
            ev = gens[chip](n, seed=7) if chip in gens else tg.synthetic_events(chip, n, seed=7)
            ev = np.ascontiguousarray(ev, dtype=np.uint32)
            plain = np.full(height * w, 0xFFFFFFFF, np.uint32)
            assert host.hostcheck_alu_rows(cid, p(ev), ctypes.c_size_t(n), ctypes.c_size_t(height), p(plain)) == 0
            for col_major in (0, 1):
                out = np.full(height * w, 0xFFFFFFFF, np.uint32)
                assert host.hostcheck_alu_rows_cta(cid, p(ev), ctypes.c_size_t(n), ctypes.c_size_t(height), p(out), col_major) == 0
                got = out.reshape(w, height).T if col_major else out.reshape(height, w)
                assert np.array_equal(got, plain.reshape(height, w)), (chip, height, col_major)


def test_cta_phases_of_the_earlier_chips(oracle, host):
    """The same walk for chips of the earlier rounds: one with four events per row, the two widest records."""
    for chip, cid, ev, orc in (("MemoryLocal", 11, tg.synthetic_memory_local_events(1021, seed=2), oracle.memory_local_trace),
                               ("Cpu", 12, tg.synthetic_cpu_events(300, seed=2), oracle.cpu_trace),
                               ("MiscInstrs", 13, tg.synthetic_misc_events(300, seed=2), oracle.misc_trace)):
        for col_major in (0, 1):
            w, h = tg.width(chip), 512
            out = np.full(h * w, 0xFFFFFFFF, np.uint32)
            p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
            assert host.hostcheck_alu_rows_cta(cid, p(ev), ctypes.c_size_t(len(ev)), ctypes.c_size_t(h), p(out), col_major) == 0
            got = out.reshape(w, h).T if col_major else out.reshape(h, w)
            assert np.array_equal(got, kb.to_monty(orc(ev, h))), chip


def test_div_rem_rows_hold_the_executor_semantics(oracle):
    n = 6000
    ev = tg.synthetic_div_rem_events(n, seed=5)
    t = oracle.chip_trace("DivRem", ev, 8192)
    word = lambda c0: sum(t[:n, c0 + k].astype(np.uint64) << np.uint64(8 * k) for k in range(4)).astype(np.uint32)
    O = tg.ALL_OPCODES
    op, hi, lo, b, c = ev[:, 4], ev[:, 5], ev[:, 6], ev[:, 7], ev[:, 8]
    assert (t[:n, 57:61].sum(axis=1) == 1).all() and (t[n:] == 0).all()
    # quotient / remainder are what the executor wrote to LO / HI
    assert np.array_equal(word(10), lo) and np.array_equal(word(14), hi)
    # c * quotient + remainder = b over 64 bits (sign-extended for DIV / MOD), when there is no division by zero / overflow
    sgn = (op == O["DIV"]) | (op == O["MOD"])
    ctq = sum(t[:n, 30 + k].astype(np.uint64) << np.uint64(8 * k) for k in range(8))
    ok = (c != 0) & ~(sgn & (b == 0x80000000) & (c == 0xFFFFFFFF))
    rem64 = np.where(sgn, hi.astype(np.int32).astype(np.int64).astype(np.uint64), hi.astype(np.uint64))
    b64 = np.where(sgn, b.astype(np.int32).astype(np.int64).astype(np.uint64), b.astype(np.uint64))
    assert ok.sum() > n // 2 and np.array_equal((ctq + rem64)[ok], b64[ok])
    # |remainder| < |c| and the sign of the remainder follows b
    assert (word(18)[ok].astype(np.uint64) < word(26)[ok].astype(np.uint64)).all()
    m = ok & sgn & (hi != 0)
    assert np.array_equal(t[:n, 88][m], t[:n, 87][m])
    # division by zero: quotient all ones, remainder b, multiplicity 0;  INT_MIN / -1: quotient INT_MIN, remainder 0
    z = c == 0
    assert z.any() and (lo[z] == 0xFFFFFFFF).all() and np.array_equal(hi[z], b[z]) and (t[:n, 90][z] == 0).all() and (t[:n, 56][z] == 1).all()
    ov = sgn & (b == 0x80000000) & (c == 0xFFFFFFFF)
    assert ov.any() and (t[:n, 61][ov] == 1).all() and (lo[ov] == 0x80000000).all() and (hi[ov] == 0).all() and (t[:n, 61][~ov] == 0).all()
    # the HI write only for DIV / DIVU
    is_div = (op == O["DIV"]) | (op == O["DIVU"])
    assert (t[:n, 91:106][~is_div] == 0).all() and np.array_equal(word(95)[is_div], hi[is_div])


def test_div_rem_rows_satisfy_the_restated_air_where_the_cpp_twin_does_not(oracle):
    """DivRemChip::eval restated as data (ziren_b200/synthetic.py _div_rem_chip): rows from trace generation - divisions by zero
    and INT_MIN / -1 among them, the rows the reference's C++ twin cannot pin - are accepted by the restated prover and verifier;
    single-cell corruptions are rejected, and so is what div_rem.hpp writes for c = 0 (quotient INT32_MAX, abs_c = 1): the Rust
    generate_trace is the one the constraints accept."""
    from ziren_b200 import synthetic
    ev = tg.synthetic_div_rem_events(500, seed=3)
    rows = oracle.chip_trace("DivRem", ev, 512)
    b, c = ev[:, 7], ev[:, 8]
    assert (c == 0).sum() >= 10 and ((b == 0x80000000) & (c == 0xFFFFFFFF)).sum() >= 3
    case = synthetic.chips_case({"DivRem": rows})
    chip = case.machine.chip("DivRem")
    assert chip.log_quotient_degree == 1 and chip.num_constraints == 120
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    proof, _ = om.prove_shard(case.traces, case.public_values)
    ok, err = om.verify_shard(proof)
    assert ok, err
    zero = int(np.argmax(c == 0))
    ovf = int(np.argmax((b == 0x80000000) & (c == 0xFFFFFFFF) & ((ev[:, 4] == 5) | (ev[:, 4] == 7))))
    for row, col in ((zero, 10), (zero, 26), (4, 14), (7, 38), (ovf, 61), (20, 90), (500, 57)):
        bad = rows.copy()
        bad[row, col] = (int(bad[row, col]) + 1) % kb.P
        p2, _ = om.prove_shard({**case.traces, "DivRem": bad}, case.public_values)
        assert not om.verify_shard(p2)[0], (row, col)
    twin = rows.copy()
    twin[zero, 10:14] = [0xFF, 0xFF, 0xFF, 0x7F]          # INT32_MAX as the quotient of a division by zero
    p2, _ = om.prove_shard({**case.traces, "DivRem": twin}, case.public_values)
    assert not om.verify_shard(p2)[0]
    twin = rows.copy()
    twin[zero, 22] = 1                                     # abs_c = max(1, |c|)
    p2, _ = om.prove_shard({**case.traces, "DivRem": twin}, case.public_values)
    assert not om.verify_shard(p2)[0]


def test_syscall_instrs_rows_satisfy_the_restated_air(oracle):
    """SyscallInstrsChip::eval restated as data (ziren_b200/synthetic.py _syscall_instrs_chip: 109 constraints - the six id tests,
    the linux / send-to-table flags, both KoalaBear range checks, the COMMIT digest words and the HALT exit code against the
    public values): rows from well-formed events are accepted, single-cell corruptions and a wrong exit code rejected."""
    from ziren_b200 import synthetic
    ev = tg.synthetic_syscall_events(400, seed=4, kind="instrs")
    rows = oracle.chip_trace("SyscallInstrs", ev, 512)
    digest, deferred, exit_code = tg.syscall_public_values(4)
    case = synthetic.syscall_instrs_case(rows, digest, deferred, exit_code)
    assert case.machine.chip("SyscallInstrs").num_constraints == 109
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    proof, _ = om.prove_shard(case.traces, case.public_values)
    ok, err = om.verify_shard(proof)
    assert ok, err
    sid = ev[:, 7] & 0xFFFF
    halt, commit, deferred_row, enter = (int(np.argmax(sid == v)) for v in (0, 0x10, 0x1A, 0x03))
    send = int(np.argmax(((ev[:, 7] >> 16) & 0xFF) == 1))
    # is_halt, next_pc of a halt, the digest word of a COMMIT, the bitmap, the deferred digest, syscall_id of ENTER_UNCONSTRAINED,
    # a range-check bit of a row sent to the table, num_extra_cycles, is_real of a padding row
    for row, col in ((halt, 5), (halt, 1), (commit, 19), (commit, 40), (deferred_row, 18), (enter, 9), (send, 47), (send, 4), (450, 76)):
        bad = rows.copy()
        bad[row, col] = (int(bad[row, col]) + 1) % kb.P
        p2, _ = om.prove_shard({"SyscallInstrs": bad}, case.public_values)
        assert not om.verify_shard(p2)[0], (row, col)
    wrong = synthetic.syscall_instrs_case(rows, digest, deferred, exit_code + 1)
    om2 = oracle.OracleMachine(wrong.machine)
    om2.setup({})
    p2, _ = om2.prove_shard(wrong.traces, wrong.public_values)
    assert not om2.verify_shard(p2)[0]


def test_syscall_rows_hold_the_chip_semantics(oracle):
    n = 4000
    ev = tg.synthetic_syscall_events(n, seed=6, kind="instrs")
    t = oracle.chip_trace("SyscallInstrs", ev, 4096)
    prev = ev[:, 7]
    sid = prev & 0xFFFF
    # each IsZeroOperation: result = (id == code), inverse * (id - code) = 1 otherwise
    for col, code in ((26, 0x03), (28, 0xF0), (30, 0x00), (32, 4246), (34, 0x10), (36, 0x1A)):
        assert np.array_equal(t[:n, col + 1] == 1, sid == code)
        d = (sid.astype(np.int64) - code) % kb.P
        assert ((t[:n, col].astype(np.uint64) * d.astype(np.uint64)) % kb.P == (sid != code)).all()
    assert np.array_equal(t[:n, 5] == 1, (sid == 0) | (sid == 4246))
    commits = (sid == 0x10) | (sid == 0x1A)
    assert commits.any() and (t[:n, 38:46].sum(axis=1) == commits).all() and (t[:n, 38:46][commits].argmax(axis=1) == ev[commits, 12]).all()
    send = (((prev >> 8) & 0xFF) != 0) | (((prev >> 16) & 0xFF) == 1)
    assert np.array_equal(t[:n, 74] == 1, send | (t[:n, 5] == 1)) and np.array_equal(t[:n, 75] == 1, send | (sid == 0x1A))
    assert (t[:n, 46:60][t[:n, 74] == 0] == 0).all() and (t[:n, 60:74][t[:n, 75] == 0] == 0).all()
    top = sum(t[:n, 46 + k].astype(np.uint32) << k for k in range(8))
    assert np.array_equal(top[t[:n, 74] == 1], (ev[:, 12] >> 24)[t[:n, 74] == 1])
    # SyscallCore: the half-words recombine, the result only for the linux calls
    ev = tg.synthetic_syscall_events(n, seed=7, kind="core")
    t = oracle.chip_trace("SyscallCore", ev, 4096)
    assert np.array_equal(t[:n, 3] + (t[:n, 4] << 16), ev[:, 12]) and np.array_equal(t[:n, 5] + (t[:n, 6] << 16), ev[:, 13])
    linux = ((ev[:, 7] >> 8) & 0xFF) != 0
    assert linux.any() and (~linux).any() and np.array_equal(t[:n, 9] == 1, linux)
    assert np.array_equal((t[:n, 7] + (t[:n, 8] << 16))[linux], ev[linux, 4]) and (t[:n, 7:9][~linux] == 0).all()


# ---- MemoryGlobalInit / MemoryGlobalFinalize --------------------------------------------------------------------------------

def _mg_case(n, seed, previous_addr):
    ev = tg.synthetic_memory_global_events(n, seed=seed)
    srt = ev[np.argsort(ev[:, 0])]
    if previous_addr == "below" and n:
        previous_addr = int(srt[0, 0]) - 1
    elif previous_addr == "below":
        previous_addr = 5
    return ev, srt, int(previous_addr)


def test_memory_global_matches_reference_golden_and_live(oracle, host):
    ev, rows = np.array(GOLD["MemoryGlobalInit"]["events"], np.uint32), np.array(GOLD["MemoryGlobalInit"]["rows"], np.uint32)
    orc = kb.to_monty(oracle.memory_global_trace(ev, 0, 64))
    assert np.array_equal(orc[: len(ev)][:, MG_TWIN_COLS], rows[:, MG_TWIN_COLS])
    got = _host_rows(host, "MemoryGlobalInit", tg.memory_global_records(ev, 0), 64)
    assert np.array_equal(got, orc)
    big = tg.synthetic_memory_global_events(5000, seed=9)
    big = big[np.argsort(big[:, 0])]
    for chip in ("MemoryGlobalInit", "MemoryGlobalFinalize"):
        ref = oracle.ref_chip_rows(chip, big, 4)
        if ref is None:
            pytest.skip("oracle/_ref not built (no reference checkout on this machine)")
        assert np.array_equal(kb.to_monty(oracle.memory_global_trace(big, 0, 8192))[:5000][:, MG_TWIN_COLS], ref[:, MG_TWIN_COLS])


@pytest.mark.parametrize("chip", ["MemoryGlobalInit", "MemoryGlobalFinalize"])
@pytest.mark.parametrize("n,height,previous_addr", [(0, 16, 0), (1, 16, 0), (1, 16, "below"), (64, 64, "below"), (65, 128, 0),
                                                    (1000, 1024, "below"), (129, 192, 0)])
def test_memory_global_product_matches_oracle_through_the_cta_phases(oracle, host, chip, n, height, previous_addr):
    ev, srt, prev = _mg_case(n, 50 + n, previous_addr)
    want = kb.to_monty(oracle.memory_global_trace(srt, prev, height))
    rec = tg.memory_global_records(ev, prev)                  # sorts
    assert np.array_equal(_host_rows(host, chip, rec, height), want)
    assert np.array_equal(_host_rows(host, chip, rec, height, cta=True), want)
    assert np.array_equal(_host_rows(host, chip, rec, height, cta=True, col_major=True), want)


def test_memory_global_rows_hold_the_chip_semantics(oracle):
    n = 3000
    ev, srt, prev = _mg_case(n, 8, "below")
    t = oracle.memory_global_trace(srt, prev, 4096)
    bits = lambda c0: sum(t[:n, c0 + k].astype(np.uint64) << np.uint64(k) for k in range(32)).astype(np.uint32)
    assert np.array_equal(bits(35), srt[:, 0]) and np.array_equal(bits(73), srt[:, 1]) and np.array_equal(t[:n, 2], srt[:, 0])
    # exactly one comparison flag per row, at the top bit where the previous address and this one differ, set in this one
    before = np.r_[np.uint32(prev), srt[:-1, 0]]
    assert (t[:n, 3:35].sum(axis=1) == 1).all()
    k = t[:n, 3:35].argmax(axis=1).astype(np.uint32)
    assert ((srt[:, 0] >> k) & 1 == 1).all() and ((before >> k) & 1 == 0).all()
    assert np.array_equal(srt[:, 0].astype(np.uint64) >> (k + 1).astype(np.uint64), before.astype(np.uint64) >> (k + 1).astype(np.uint64))
    assert t[0, 106] == 0 and (t[1:n, 106] == 1).all() and t[0, 109] == 1 and (t[1:n, 109] == 0).all()
    assert t[n - 1, 110] == 1 and (t[: n - 1, 110] == 0).all() and (t[n:] == 0).all()
    assert int(t[0, 107]) * prev % kb.P == 1 and t[0, 108] == 0
    # a first shard: previous address 0, no comparison in the first row
    t0 = oracle.memory_global_trace(srt, 0, 4096)
    assert (t0[0, 3:35] == 0).all() and t0[0, 107] == 0 and t0[0, 108] == 1 and t0[0, 109] == 0
    # the running products of the top byte's bits
    top = srt[:, 0] >> 24
    for j, col in enumerate(range(67, 73)):
        mask = (1 << (j + 2)) - 1
        assert np.array_equal(t[:n, col] == 1, (top & mask) == mask)


def test_oracle_rejects_unsorted_memory_events(oracle):
    ev = tg.synthetic_memory_global_events(40, seed=1)
    srt = ev[np.argsort(ev[:, 0])]
    with pytest.raises(RuntimeError):
        oracle.memory_global_trace(srt[::-1].copy(), 0, 64)
    with pytest.raises(RuntimeError):
        oracle.memory_global_trace(srt, int(srt[0, 0]) + 1, 64)


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from ziren_b200 import synthetic as syn
    from ziren_b200.prover import B200Prover
    prover = B200Prover(syn.mini_case().machine, device=0)
    yield torch, prover
    prover.close()


@pytest.mark.gpu
@pytest.mark.parametrize("chip", ROW_CHIPS + ("MemoryGlobalInit", "MemoryGlobalFinalize"))
@pytest.mark.parametrize("n,log_h,col_major,on_device", [(5000, 13, False, False), (5000, 13, True, True), (129, 8, True, False),
                                                         (1, 4, False, True), (0, 4, True, False)])
def test_gpu_trace_matches_oracle(gpu, oracle, chip, n, log_h, col_major, on_device):
    torch, prover = gpu
    w, h = tg.width(chip), 1 << log_h
    if chip.startswith("MemoryGlobal"):
        raw, srt, prev = _mg_case(n, 20 + n, "below" if n % 2 else 0)
        ev = tg.memory_global_records(raw, prev)
        want = kb.to_monty(oracle.memory_global_trace(srt, prev, h))
    else:
        ev = synthetic(chip, n, 20 + n)
        want = kb.to_monty(oracle.chip_trace(chip, ev, h))
    out = torch.full((h * w,), -1, dtype=torch.int32, device="cuda")
    src = torch.from_numpy(ev.view(np.int32)).cuda() if on_device and n else ev
    prover.generate_alu_trace(chip, src, log_h, out, col_major=col_major)
    got = out.cpu().numpy().view(np.uint32)
    got = got.reshape(w, h).T if col_major else got.reshape(h, w)
    assert np.array_equal(got, want)
    if n >= 96 and chip in GOLD and not chip.startswith("MemoryGlobal"):
        gev, grows = np.array(GOLD[chip]["events"], np.uint32), np.array(GOLD[chip]["rows"], np.uint32)
        out2 = torch.zeros((128 * w,), dtype=torch.int32, device="cuda")
        prover.generate_alu_trace(chip, gev, 7, out2)
        assert np.array_equal(out2.cpu().numpy().view(np.uint32).reshape(128, w)[: len(gev)], grows)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["events_host", "events_device"])
def test_div_rem_shard_proves_from_event_records(gpu, oracle, mode):
    """The DivRem table under its restated constraints, handed to zkb200_commit as CompAluEvent records: the proof is the
    oracle's proof over the oracle's rows, word for word."""
    from ziren_b200 import synthetic
    from ziren_b200.prover import B200Prover, EventTrace
    torch, _ = gpu
    ev = tg.synthetic_div_rem_events(3000, seed=33)
    log_h = tg.padded_log_height(len(ev))
    case = synthetic.chips_case({"DivRem": oracle.chip_trace("DivRem", ev, 1 << log_h)})
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    want, _ = om.prove_shard(case.traces, case.public_values)
    prover = B200Prover(case.machine, device=0)
    try:
        pk = prover.setup({})
        inputs = {k: kb.to_monty(v) for k, v in case.traces.items() if k != "DivRem"}
        d = torch.from_numpy(ev.view(np.int32)).cuda() if mode == "events_device" else ev
        inputs["DivRem"] = EventTrace(d, log_h, tg.width("DivRem"))
        got, _ = prover.prove_shard(pk, inputs, case.public_values)
        ok, err = om.verify_shard(got)
        assert ok, err
        assert np.array_equal(got, want)
        pk.free()
    finally:
        prover.close()


@pytest.mark.gpu
def test_syscall_instrs_shard_proves_from_event_records(gpu, oracle):
    """The SyscallInstrs table under its restated constraints (public values: digests and exit code), handed to zkb200_commit as
    the record's SyscallEvent vector."""
    from ziren_b200 import synthetic
    from ziren_b200.prover import B200Prover, EventTrace
    ev = tg.synthetic_syscall_events(1800, seed=6, kind="instrs")
    log_h = tg.padded_log_height(len(ev))
    case = synthetic.syscall_instrs_case(oracle.chip_trace("SyscallInstrs", ev, 1 << log_h), *tg.syscall_public_values(6))
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    want, _ = om.prove_shard(case.traces, case.public_values)
    prover = B200Prover(case.machine, device=0)
    try:
        pk = prover.setup({})
        got, _ = prover.prove_shard(pk, {"SyscallInstrs": EventTrace(ev, log_h, tg.width("SyscallInstrs"))}, case.public_values)
        ok, err = om.verify_shard(got)
        assert ok, err
        assert np.array_equal(got, want)
        pk.free()
    finally:
        prover.close()

"""Trace generation of the MemoryInstrs chip: MemInstrEvent records of sixteen words.

CPU tests: the oracle (oracle/tracegen.h mem_instr_row) and the product's row filler compiled for the host
(ziren_b200/csrc/tracegen.cuh fill_mem_instr) against golden rows written by the REFERENCE'S OWN C++
(crates/core/machine/include/memory_instrs.hpp; tests/golden/mem_instr_rows.json) and, when oracle/_ref is present, against
that C++ live; the loaded value against the executor's semantics.  GPU: the CUDA kernel through the C ABI."""
import ctypes
import json
import os

import numpy as np
import pytest

from ziren_b200 import field as kb
from ziren_b200 import tracegen as tg

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "mem_instr_rows.json")))
CHIP = "MemoryInstrs"
CHIP_ID = 10        # AluChip::ALU_MEMINSTR, csrc/tracegen.cuh


def _host_rows(host, ev, height):
    ev = np.ascontiguousarray(ev, dtype=np.uint32).reshape(-1, tg.COMP_EVENT_WORDS)
    out = np.zeros((height, tg.width(CHIP)), np.uint32)
    rc = host.hostcheck_alu_rows(CHIP_ID, ev.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(ev)), ctypes.c_size_t(height),
                                 out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def test_oracle_and_product_match_reference_golden_rows(oracle, host):
    ev, rows = np.array(GOLD["events"], np.uint32), np.array(GOLD["rows"], np.uint32)
    assert GOLD["width"] == tg.width(CHIP) == oracle.MEMINSTR_WIDTH == host.hostcheck_alu_width(CHIP_ID) == 79
    assert tg.event_words(CHIP) == 16
    assert np.array_equal(kb.to_monty(oracle.mem_instr_trace(ev, 128))[: len(ev)], rows)
    assert np.array_equal(_host_rows(host, ev, 128)[: len(ev)], rows)


def test_oracle_and_product_match_reference_cpp_live(oracle, host):
    ev = tg.synthetic_mem_instr_events(6000, seed=3)
    ref = oracle.ref_mem_instr_rows(ev)
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference checkout on this machine)")
    orc = kb.to_monty(oracle.mem_instr_trace(ev, 8192))
    assert np.array_equal(orc[: len(ev)], ref)
    assert np.array_equal(_host_rows(host, ev, 8192), orc)          # zero padding rows included


def test_unused_record_word_and_padding_bytes_are_ignored(oracle, host):
    """Word 14 of a read record and the three bytes above the opcode are padding in the Rust struct: any value there leaves
    the rows unchanged."""
    ev = tg.synthetic_mem_instr_events(500, seed=9)
    want = _host_rows(host, ev, 512)
    noisy = ev.copy()
    reads = noisy[:, 8] == 0
    noisy[reads, 14] = 0xDEADBEEF
    noisy[:, 4] |= np.uint32(0xABCDEF00)
    assert np.array_equal(_host_rows(host, noisy, 512), want)
    assert np.array_equal(kb.to_monty(oracle.mem_instr_trace(noisy, 512)), want)


def test_rows_hold_the_executor_semantics(oracle):
    """a = sign/zero extension of unsigned_mem_val for the byte and halfword loads, = unsigned_mem_val for the word loads
    (execute_load, executor.rs:1925-2003); one opcode flag per row; the address word is b + c."""
    n = 4000
    ev = tg.synthetic_mem_instr_events(n, seed=5)
    t = oracle.mem_instr_trace(ev, 4096)
    word = lambda c0: sum(t[:n, c0 + k].astype(np.uint64) << np.uint64(8 * k) for k in range(4)).astype(np.uint32)
    op, a = ev[:, 4], ev[:, 5]
    um, neg = word(70), t[:n, 76]
    O = tg.OPCODES
    assert (t[:n, 16:30].sum(axis=1) == 1).all() and (t[n:] == 0).all()
    assert np.array_equal(word(30), (ev[:, 6] + ev[:, 7]).astype(np.uint32))
    assert np.array_equal(t[:n, 34] + t[:n, 35], word(30))
    for o_, ext in (("LB", 0xFFFFFF00), ("LH", 0xFFFF0000)):
        m = op == O[o_]
        assert m.any() and neg[m].any() and not neg[m].all()
        assert np.array_equal(a[m], um[m] | np.where(neg[m] == 1, np.uint32(ext), np.uint32(0)))
    for o_ in ("LBU", "LHU", "LW", "LWL", "LWR", "LL"):
        m = op == O[o_]
        assert m.any() and np.array_equal(a[m], um[m]) and (neg[m] == 0).all()
    stores = op >= O["SB"]
    assert (um[stores] == 0).all() and (t[:n, 74:77][stores] == 0).all()
    # most_sig_bytes_zero: result = the address is below 256, inverse * sum = 1 otherwise
    upper = t[:n, 31].astype(np.uint64) + t[:n, 32] + t[:n, 33]
    assert np.array_equal(t[:n, 78] == 1, upper == 0) and (upper == 0).any()
    nz = upper != 0
    assert ((t[:n, 77][nz].astype(np.uint64) * upper[nz]) % kb.P == 1).all() and (t[:n, 77][~nz] == 0).all()


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from ziren_b200 import synthetic
    from ziren_b200.prover import B200Prover
    prover = B200Prover(synthetic.mini_case().machine, device=0)
    yield torch, prover
    prover.close()


@pytest.mark.gpu
@pytest.mark.parametrize("n,log_h,col_major,on_device", [(5000, 13, False, False), (5000, 13, True, True), (129, 8, True, False),
                                                         (1, 4, False, True), (0, 4, True, False)])
def test_gpu_mem_instr_trace_matches_oracle(gpu, oracle, n, log_h, col_major, on_device):
    torch, prover = gpu
    ev = tg.synthetic_mem_instr_events(n, seed=20 + n)
    w, h = tg.width(CHIP), 1 << log_h
    out = torch.full((h * w,), -1, dtype=torch.int32, device="cuda")
    src = torch.from_numpy(ev.view(np.int32)).cuda() if on_device and n else ev
    prover.generate_alu_trace(CHIP, src, log_h, out, col_major=col_major)
    got = out.cpu().numpy().view(np.uint32)
    got = got.reshape(w, h).T if col_major else got.reshape(h, w)
    assert np.array_equal(got, kb.to_monty(oracle.mem_instr_trace(ev, h)))
    if n >= 96:
        gev, grows = np.array(GOLD["events"], np.uint32), np.array(GOLD["rows"], np.uint32)
        out2 = torch.zeros((128 * w,), dtype=torch.int32, device="cuda")
        prover.generate_alu_trace(CHIP, gev, 7, out2)
        assert np.array_equal(out2.cpu().numpy().view(np.uint32).reshape(128, w)[: len(gev)], grows)

"""Trace generation of the MiscInstrs chip: fifteen-word MiscEvent records, 72 columns of which
44 are a union read per opcode family (SEXT / TEQ, EXT, INS, MADDU / MSUBU / MADD / MSUB).

CPU tests: the oracle (oracle/tracegen.h misc_row) and the product's row filler compiled for the host
(ziren_b200/csrc/tracegen.cuh fill_misc) against golden rows written by the REFERENCE'S OWN C++
(crates/core/machine/include/misc_instrs.hpp; tests/golden/misc_rows.json) and, when oracle/_ref is present, against that
C++ live; the rows against the executor's semantics.  GPU: the CUDA kernel through the C ABI."""
import ctypes
import json
import os

import numpy as np
import pytest

from ziren_b200 import field as kb
from ziren_b200 import tracegen as tg

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "misc_rows.json")))
CHIP = "MiscInstrs"
CHIP_ID = 13        # AluChip::ALU_MISC, csrc/tracegen.cuh


def _host_rows(host, ev, height):
    ev = np.ascontiguousarray(ev, dtype=np.uint32).reshape(-1, tg.MISC_EVENT_WORDS)
    out = np.full((height, tg.width(CHIP)), 0xFFFFFFFF, np.uint32)
    rc = host.hostcheck_alu_rows(CHIP_ID, ev.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(ev)), ctypes.c_size_t(height),
                                 out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def test_oracle_and_product_match_reference_golden_rows(oracle, host):
    ev, rows = np.array(GOLD["events"], np.uint32), np.array(GOLD["rows"], np.uint32)
    assert GOLD["width"] == tg.width(CHIP) == oracle.MISC_WIDTH == host.hostcheck_alu_width(CHIP_ID) == 72
    assert tg.event_words(CHIP) == oracle.MISC_EVENT_WORDS == 15
    assert set(ev[:, 4]) == {tg.ALL_OPCODES[o] for o in tg.MISC_OPCODES}
    assert np.array_equal(kb.to_monty(oracle.misc_trace(ev, 128))[: len(ev)], rows)
    assert np.array_equal(_host_rows(host, ev, 128)[: len(ev)], rows)


def test_oracle_and_product_match_reference_cpp_live(oracle, host):
    ev = tg.synthetic_misc_events(9000, seed=3)
    ref = oracle.ref_misc_rows(ev)
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference checkout on this machine)")
    orc = kb.to_monty(oracle.misc_trace(ev, 16384))
    assert np.array_equal(orc[: len(ev)], ref)
    assert np.array_equal(_host_rows(host, ev, 16384), orc)         # zero padding rows included


def test_rows_hold_the_executor_semantics(oracle):
    """EXT: a << lsb ... the columns recombine to what execute_ext / execute_ins / execute_madd* computed."""
    n = 6000
    ev = tg.synthetic_misc_events(n, seed=5)
    t = oracle.misc_trace(ev, 8192)
    word = lambda c0: sum(t[:n, c0 + k].astype(np.uint64) << np.uint64(8 * k) for k in range(4)).astype(np.uint32)
    O = tg.ALL_OPCODES
    op, a, b, c, prev_a = ev[:, 4], ev[:, 5], ev[:, 6], ev[:, 7], ev[:, 8]
    assert (t[:n, 64:72].sum(axis=1) == 1).all() and (t[n:] == 0).all()
    # EXT: sll_val >> (31 - msbd) = a
    m = op == O["EXT"]
    assert m.any() and np.array_equal(word(22)[m] >> (np.uint32(31) - (c[m] >> 5)), a[m])
    # INS: add_val rotated left by (msb + 1) = a  (trace.rs:236-262 and the AIR's final rotate)
    m = op == O["INS"]
    msb = (c[m] >> 5).astype(np.uint64)
    add_val = word(38)[m].astype(np.uint64)
    rot = (31 - msb).astype(np.uint64)            # ror by 31 - msb
    got = ((add_val >> rot) | (add_val << (np.uint64(32) - rot))) & np.uint64(0xFFFFFFFF)
    got = np.where(rot == 0, add_val, got)
    assert m.any() and np.array_equal(got.astype(np.uint32), a[m])
    # MADD family: add_operation.value = multiply + src2 (low word), and for the additive ones it is a
    for name in ("MADDU", "MADD"):
        m = op == O[name]
        assert m.any() and np.array_equal(word(28)[m], a[m]) and np.array_equal(word(32)[m], ev[m, 9])
    for name in ("MSUBU", "MSUB"):               # multiply + result = previous accumulator
        m = op == O[name]
        assert m.any() and np.array_equal(word(28)[m], prev_a[m]) and np.array_equal(word(32)[m], ev[m, 12])
    # SEXT: a = sign extension of the byte / halfword of b; most_sig_bit is the sign
    m = op == O["SEXT"]
    seh = t[:n, 34][m] == 1
    neg = t[:n, 20][m] == 1
    want = np.where(seh, np.where(neg, b[m] | np.uint32(0xFFFF0000), b[m] & np.uint32(0xFFFF)),
                    np.where(neg, b[m] | np.uint32(0xFFFFFF00), b[m] & np.uint32(0xFF)))
    assert m.any() and seh.any() and (~seh).any() and neg.any() and np.array_equal(a[m], want)
    # TEQ: a != b, so a_eq_b.result = 0, with some equal bytes
    m = op == O["TEQ"]
    assert m.any() and (t[:n, 32][m] == 0).all() and (t[:n, 23:30:2][m] == 1).any()


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
def test_gpu_misc_trace_matches_oracle(gpu, oracle, n, log_h, col_major, on_device):
    torch, prover = gpu
    ev = tg.synthetic_misc_events(n, seed=20 + n)
    w, h = tg.width(CHIP), 1 << log_h
    out = torch.full((h * w,), -1, dtype=torch.int32, device="cuda")
    src = torch.from_numpy(ev.view(np.int32)).cuda() if on_device and n else ev
    prover.generate_alu_trace(CHIP, src, log_h, out, col_major=col_major)
    got = out.cpu().numpy().view(np.uint32)
    got = got.reshape(w, h).T if col_major else got.reshape(h, w)
    assert np.array_equal(got, kb.to_monty(oracle.misc_trace(ev, h)))
    if n >= 96:
        gev, grows = np.array(GOLD["events"], np.uint32), np.array(GOLD["rows"], np.uint32)
        out2 = torch.zeros((128 * w,), dtype=torch.int32, device="cuda")
        prover.generate_alu_trace(CHIP, gev, 7, out2)
        assert np.array_equal(out2.cpu().numpy().view(np.uint32).reshape(128, w)[: len(gev)], grows)

"""Trace generation of the Cpu chip, the tallest table of a core shard: one 28-word
`zkb200_cpu_event` (the flattened CpuEvent + Instruction, include/zkb200.h) per row.

CPU tests: the oracle (oracle/tracegen.h cpu_row) and the product's row filler compiled for the host
(ziren_b200/csrc/tracegen.cuh fill_cpu) against golden rows written by the REFERENCE'S OWN C++
(crates/core/machine/include/cpu.hpp; tests/golden/cpu_rows.json, the first event is the reference's own test vector,
cpu/trace.rs:283-306) and, when oracle/_ref is present, against that C++ live.  GPU: the CUDA kernel through the C ABI."""
import ctypes
import json
import os

import numpy as np
import pytest

from ziren_b200 import field as kb
from ziren_b200 import tracegen as tg

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "cpu_rows.json")))
CHIP = "Cpu"
CHIP_ID = 12        # AluChip::ALU_CPU, csrc/tracegen.cuh


def _host_rows(host, ev, height):
    ev = np.ascontiguousarray(ev, dtype=np.uint32).reshape(-1, tg.CPU_EVENT_WORDS)
    out = np.full((height, tg.width(CHIP)), 0xFFFFFFFF, np.uint32)
    rc = host.hostcheck_alu_rows(CHIP_ID, ev.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(ev)), ctypes.c_size_t(height),
                                 out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def test_oracle_and_product_match_reference_golden_rows(oracle, host):
    ev, rows = np.array(GOLD["events"], np.uint32), np.array(GOLD["rows"], np.uint32)
    assert GOLD["width"] == tg.width(CHIP) == oracle.CPU_WIDTH == host.hostcheck_alu_width(CHIP_ID) == 67
    assert tg.event_words(CHIP) == oracle.CPU_EVENT_WORDS == 28
    assert set(ev[:, 9] & 0xFF) >= set(tg.ALL_OPCODES.values())           # every opcode is in the fixture
    assert np.array_equal(kb.to_monty(oracle.cpu_trace(ev, 128))[: len(ev)], rows)
    assert np.array_equal(_host_rows(host, ev, 128)[: len(ev)], rows)
    # the reference's own test row: the b / c value words come from the records (5), not from the event's b / c (10, 15)
    t = oracle.cpu_trace(ev[:1], 16)
    assert list(t[0, 47:51]) == [5, 0, 0, 0] and list(t[0, 56:60]) == [5, 0, 0, 0] and list(t[0, 30:34]) == [1, 0, 0, 0]
    assert t[0, 8] == 0 and t[0, 9] == 29 and t[0, 19] == 0 and t[0, 20] == 1 and t[0, 25] == 1


def test_oracle_and_product_match_reference_cpp_live(oracle, host):
    ev = tg.synthetic_cpu_events(9000, seed=3)
    ref = oracle.ref_cpu_rows(ev)
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference checkout on this machine)")
    orc = kb.to_monty(oracle.cpu_trace(ev, 16384))
    assert np.array_equal(orc[: len(ev)], ref)
    assert np.array_equal(_host_rows(host, ev, 16384), orc)         # padding rows included


def test_padding_rows_and_flags(oracle):
    """Padding rows are zero except imm_b = imm_c = is_rw_a = 1 (cpu/trace.rs:60-63); HALT and SYS_EXT_GROUP end sequential
    flow; shard_to_send / clk_to_send only on the rows whose memory access another chip checks."""
    n = 6000
    ev = tg.synthetic_cpu_events(n, seed=5)
    t = oracle.cpu_trace(ev, 8192)
    pad = np.zeros(67, np.uint32)
    pad[[19, 20, 22]] = 1
    assert (t[n:] == pad).all()
    op = ev[:, 9] & 0xFF
    O = tg.ALL_OPCODES
    sys_rows = op == O["SYSCALL"]
    code = ev[:, 15] & 0xFFFF
    halt = sys_rows & ((code == 0) | (code == 4246))
    assert halt.any() and np.array_equal(t[:n, 24] == 1, halt)
    assert np.array_equal(t[:n, 21], np.where(sys_rows, ev[:, 15] >> 24, 0))
    flow = ((op >= O["BEQ"]) & (op <= O["JumpDirect"])) | halt
    assert np.array_equal(t[:n, 25] == 1, ~flow)
    send = t[:n, 23] == 1
    assert np.array_equal(t[:n, 3], np.where(send, 3, 0)) and np.array_equal(t[:n, 4], np.where(send, ev[:, 0], 0))
    assert np.array_equal(t[:n, 1] + (t[:n, 2] << 16), ev[:, 0]) and (t[:n, 2] > 0).any()
    assert (t[:n, 65] == 1).all()


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
def test_gpu_cpu_trace_matches_oracle(gpu, oracle, n, log_h, col_major, on_device):
    torch, prover = gpu
    ev = tg.synthetic_cpu_events(n, seed=20 + n)
    w, h = tg.width(CHIP), 1 << log_h
    out = torch.full((h * w,), -1, dtype=torch.int32, device="cuda")
    src = torch.from_numpy(ev.view(np.int32)).cuda() if on_device and n else ev
    prover.generate_alu_trace(CHIP, src, log_h, out, col_major=col_major)
    got = out.cpu().numpy().view(np.uint32)
    got = got.reshape(w, h).T if col_major else got.reshape(h, w)
    assert np.array_equal(got, kb.to_monty(oracle.cpu_trace(ev, h)))
    if n >= 96:
        gev, grows = np.array(GOLD["events"], np.uint32), np.array(GOLD["rows"], np.uint32)
        out2 = torch.zeros((128 * w,), dtype=torch.int32, device="cuda")
        prover.generate_alu_trace(CHIP, gev, 7, out2)
        assert np.array_equal(out2.cpu().numpy().view(np.uint32).reshape(128, w)[: len(gev)], grows)

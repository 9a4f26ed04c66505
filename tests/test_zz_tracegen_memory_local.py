"""Trace generation of the MemoryLocal chip: seven-word MemoryLocalEvent records, FOUR per row.

CPU tests: the oracle (oracle/tracegen.h memory_local_trace) and the product's row filler compiled for the host
(ziren_b200/csrc/tracegen.cuh fill_memory_local) against golden entries written by the REFERENCE'S OWN C++
(crates/core/machine/include/memory_local.hpp; tests/golden/memory_local_entries.json) laid out four to a row as
MemoryLocalChip::generate_trace does (memory/local.rs:146-190) and, when oracle/_ref is present, against that C++ live.
GPU: the CUDA kernel through the C ABI."""
import ctypes
import json
import os

import numpy as np
import pytest

from ziren_b200 import field as kb
from ziren_b200 import tracegen as tg

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "memory_local_entries.json")))
CHIP = "MemoryLocal"
CHIP_ID = 11        # AluChip::ALU_MEMLOCAL, csrc/tracegen.cuh


def _rows_from_entries(entries, height):
    """generate_trace's layout: event 4 i + k is entry k of row i, everything past the last event zero."""
    n = len(entries)
    flat = np.zeros((height * 4, 14), np.uint32)
    flat[:n] = entries
    return flat.reshape(height, 56)


def _host_rows(host, ev, height):
    ev = np.ascontiguousarray(ev, dtype=np.uint32).reshape(-1, 7)
    out = np.full((height, tg.width(CHIP)), 0xFFFFFFFF, np.uint32)
    rc = host.hostcheck_alu_rows(CHIP_ID, ev.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(ev)), ctypes.c_size_t(height),
                                 out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def test_oracle_and_product_match_reference_golden_entries(oracle, host):
    ev, entries = np.array(GOLD["events"], np.uint32), np.array(GOLD["entries"], np.uint32)
    assert GOLD["entry_width"] == oracle.MEMLOCAL_ENTRY_WIDTH == 14
    assert tg.width(CHIP) == oracle.MEMLOCAL_WIDTH == host.hostcheck_alu_width(CHIP_ID) == 56
    assert tg.events_per_row(CHIP) == 4 and tg.event_words(CHIP) == 7
    assert len(ev) % 4 == 2                                    # the last row is half full
    want = _rows_from_entries(entries, 32)
    assert np.array_equal(kb.to_monty(oracle.memory_local_trace(ev, 32)), want)
    assert np.array_equal(_host_rows(host, ev, 32), want)


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 4096, 4099])
def test_oracle_and_product_match_reference_cpp_live(oracle, host, n):
    ev = tg.synthetic_memory_local_events(n, seed=3 + n)
    ref = oracle.ref_memory_local_entries(ev) if n else np.zeros((0, 14), np.uint32)
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference checkout on this machine)")
    h = 1 << tg.padded_log_height(n, chip=CHIP)
    want = _rows_from_entries(ref, h)
    assert np.array_equal(kb.to_monty(oracle.memory_local_trace(ev, h)), want)
    assert np.array_equal(_host_rows(host, ev, h), want)


def test_more_events_than_rows_is_an_error(oracle, host):
    ev = tg.synthetic_memory_local_events(65, seed=1)
    with pytest.raises(RuntimeError, match="more events than rows"):
        oracle.memory_local_trace(ev, 16)
    out = np.zeros((16, 56), np.uint32)
    assert host.hostcheck_alu_rows(CHIP_ID, ev.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(65), ctypes.c_size_t(16),
                                   out.ctypes.data_as(ctypes.c_void_p)) == 1
    assert oracle.memory_local_trace(ev[:64], 16).shape == (16, 56)     # exactly full


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
@pytest.mark.parametrize("n,log_h,col_major,on_device", [(20000, 13, False, False), (20001, 13, True, True), (513, 8, True, False),
                                                         (1, 4, False, True), (0, 4, True, False), (64, 4, False, False)])
def test_gpu_memory_local_trace_matches_oracle(gpu, oracle, n, log_h, col_major, on_device):
    torch, prover = gpu
    ev = tg.synthetic_memory_local_events(n, seed=20 + n)
    w, h = tg.width(CHIP), 1 << log_h
    out = torch.full((h * w,), -1, dtype=torch.int32, device="cuda")
    src = torch.from_numpy(ev.view(np.int32)).cuda() if on_device and n else ev
    prover.generate_alu_trace(CHIP, src, log_h, out, col_major=col_major)
    got = out.cpu().numpy().view(np.uint32)
    got = got.reshape(w, h).T if col_major else got.reshape(h, w)
    assert np.array_equal(got, kb.to_monty(oracle.memory_local_trace(ev, h)))
    if n >= 20000:
        gev, gent = np.array(GOLD["events"], np.uint32), np.array(GOLD["entries"], np.uint32)
        out2 = torch.zeros((32 * w,), dtype=torch.int32, device="cuda")
        prover.generate_alu_trace(CHIP, gev, 5, out2)
        assert np.array_equal(out2.cpu().numpy().view(np.uint32).reshape(32, w), _rows_from_entries(gent, 32))
        from ziren_b200.prover import ZkbError
        with pytest.raises(ZkbError, match="more events than rows"):
            prover.generate_alu_trace(CHIP, ev, 12, out)               # 20000 events need 5000 rows

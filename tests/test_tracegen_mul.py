"""Trace generation of the Mul chip: CompAluEvent records of sixteen words.

CPU tests: the oracle (oracle/tracegen.h mul_row) and the product's row filler compiled for the host
(ziren_b200/csrc/tracegen.cuh fill_mul) against golden rows written by the REFERENCE'S OWN C++
(crates/core/machine/include/mul.hpp; tests/golden/mul_rows.json, the first event is the reference's own test vector,
mul/mod.rs:546-574) and, when oracle/_ref is present, against that C++ live.  GPU: the CUDA kernel through the C ABI."""
import ctypes
import json
import os

import numpy as np
import pytest

from ziren_b200 import field as kb
from ziren_b200 import tracegen as tg

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "mul_rows.json")))
MUL_ID = 9          # AluChip::ALU_MUL, csrc/tracegen.cuh


def _host_rows(host, ev, height):
    ev = np.ascontiguousarray(ev, dtype=np.uint32).reshape(-1, tg.COMP_EVENT_WORDS)
    out = np.zeros((height, tg.width("Mul")), np.uint32)
    rc = host.hostcheck_alu_rows(MUL_ID, ev.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(ev)), ctypes.c_size_t(height),
                                 out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def test_oracle_and_product_match_reference_golden_rows(oracle, host):
    ev, rows = np.array(GOLD["events"], np.uint32), np.array(GOLD["rows"], np.uint32)
    assert GOLD["width"] == tg.width("Mul") == oracle.MUL_WIDTH == host.hostcheck_alu_width(MUL_ID) == 58
    assert np.array_equal(kb.to_monty(oracle.mul_trace(ev, 128))[: len(ev)], rows)
    assert np.array_equal(_host_rows(host, ev, 128)[: len(ev)], rows)


def test_oracle_and_product_match_reference_cpp_live(oracle, host):
    ev = tg.synthetic_mul_events(6000, seed=3)
    ref = oracle.ref_mul_rows(ev)
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference checkout on this machine)")
    orc = kb.to_monty(oracle.mul_trace(ev, 8192))
    assert np.array_equal(orc[: len(ev)], ref)
    assert np.array_equal(_host_rows(host, ev, 8192), orc)          # zero padding rows included


def test_rows_hold_the_product(oracle):
    ev = tg.synthetic_mul_events(3000, seed=5)
    t = oracle.mul_trace(ev, 4096)
    prod = sum(t[:3000, 26 + k].astype(object) << (8 * k) for k in range(8))
    b, c, op = ev[:, 7], ev[:, 8], ev[:, 4]
    signed = op == tg.OPCODES["MULT"]
    want_s = (b.astype(np.int32).astype(object) * c.astype(np.int32).astype(object)) % (1 << 64)
    want_u = b.astype(object) * c.astype(object)
    assert all(int(p) == int(ws if s else wu) for p, ws, wu, s in zip(prod, want_s, want_u, signed))
    assert (t[3000:] == 0).all() and (t[:3000, 41] == 1).all()
    # low word = a, high word = hi for MULT / MULTU
    a = sum(t[:3000, 6 + k].astype(np.uint64) << np.uint64(8 * k) for k in range(4))
    assert np.array_equal(a.astype(np.uint32), ev[:, 6])


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
def test_gpu_mul_trace_matches_oracle(gpu, oracle, n, log_h, col_major, on_device):
    torch, prover = gpu
    ev = tg.synthetic_mul_events(n, seed=20 + n)
    w, h = tg.width("Mul"), 1 << log_h
    out = torch.full((h * w,), -1, dtype=torch.int32, device="cuda")
    src = torch.from_numpy(ev.view(np.int32)).cuda() if on_device and n else ev
    prover.generate_alu_trace("Mul", src, log_h, out, col_major=col_major)
    got = out.cpu().numpy().view(np.uint32)
    got = got.reshape(w, h).T if col_major else got.reshape(h, w)
    assert np.array_equal(got, kb.to_monty(oracle.mul_trace(ev, h)))
    if n >= 96:
        gev, grows = np.array(GOLD["events"], np.uint32), np.array(GOLD["rows"], np.uint32)
        out2 = torch.zeros((128 * w,), dtype=torch.int32, device="cuda")
        prover.generate_alu_trace("Mul", gev, 7, out2)
        assert np.array_equal(out2.cpu().numpy().view(np.uint32).reshape(128, w)[: len(gev)], grows)

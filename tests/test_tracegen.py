"""Trace generation of the core ALU chips.

CPU tests: the oracle restatement (oracle/tracegen.h) and the product's row fillers compiled for the
host (ziren_b200/csrc/tracegen.cuh through tests/hostcheck) against the golden rows written by the
REFERENCE'S OWN C++ row fillers (tests/golden/alu_rows.json, tests/golden/gen_alu_golden.py) and,
when oracle/_ref is present, against that C++ live on thousands of seeded events.
GPU tests: the CUDA kernels through the C ABI against the oracle, bit-exact, in both layouts."""
import ctypes
import json
import os

import numpy as np
import pytest

from ziren_b200 import field as kb
from ziren_b200 import tracegen as tg

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "alu_rows.json")))["chips"]
CHIPS = list(tg.ALU_CHIPS)


def _host_rows(host, chip, ev, height):
    ev = np.ascontiguousarray(ev, dtype=np.uint32).reshape(-1, 7)
    out = np.zeros((height, tg.width(chip)), np.uint32)
    rc = host.hostcheck_alu_rows(CHIPS.index(chip), ev.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(ev)),
                                 ctypes.c_size_t(height), out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def test_widths_are_the_column_struct_sizes(oracle, host):
    for i, chip in enumerate(CHIPS):
        assert tg.width(chip) == GOLD[chip]["width"] == oracle.alu_width(chip) == host.hostcheck_alu_width(i)


@pytest.mark.parametrize("chip", CHIPS)
def test_oracle_matches_reference_golden_rows(oracle, chip):
    ev, rows = np.array(GOLD[chip]["events"], np.uint32), np.array(GOLD[chip]["rows"], np.uint32)
    got = oracle.alu_trace(chip, ev, 128)
    assert np.array_equal(kb.to_monty(got[: len(ev)]), rows)


@pytest.mark.parametrize("chip", CHIPS)
def test_product_row_fillers_match_reference_golden_rows(host, chip):
    ev, rows = np.array(GOLD[chip]["events"], np.uint32), np.array(GOLD[chip]["rows"], np.uint32)
    assert np.array_equal(_host_rows(host, chip, ev, 128)[: len(ev)], rows)


@pytest.mark.parametrize("chip", CHIPS)
def test_oracle_and_product_match_reference_cpp_live(oracle, host, chip):
    ev = tg.synthetic_events(chip, 6000, seed=3)
    ref = oracle.ref_alu_rows(chip, ev)
    if ref is None:
        pytest.skip("oracle/_ref not built (no reference checkout on this machine)")
    h = 1 << tg.padded_log_height(len(ev))
    orc = kb.to_monty(oracle.alu_trace(chip, ev, h))
    prod = _host_rows(host, chip, ev, h)
    assert np.array_equal(orc[: len(ev)], ref)
    assert np.array_equal(prod, orc)          # padding rows included


@pytest.mark.parametrize("chip", CHIPS)
def test_padding_rows(oracle, host, chip):
    # generate_trace pads with the chip's dummy row: zeros, except ShiftLeft / ShiftRight / CloClz
    t = oracle.alu_trace(chip, np.zeros((0, 7), np.uint32), 16)
    assert t.shape == (16, tg.width(chip)) and (t == t[0]).all()
    nz = {int(i): int(v) for i, v in enumerate(t[0]) if v}
    want = {"ShiftLeft": {22: 1, 30: 1, 39: 1}, "ShiftRight": {10: 1, 18: 1}, "CloClz": {2: 32, 14: 1}}.get(chip, {})
    assert nz == want
    assert np.array_equal(_host_rows(host, chip, np.zeros((0, 7), np.uint32), 16), kb.to_monty(t))


def test_padded_height_rule():
    # next_power_of_two: at least 16 rows, else the next power of two; a fixed shape height wins
    assert [tg.padded_log_height(n) for n in (0, 1, 15, 16, 17, 1000, 1024, 1025)] == [4, 4, 4, 4, 5, 10, 10, 11]
    assert tg.padded_log_height(5, fixed_log2_rows=12) == 12
    with pytest.raises(ValueError):
        tg.padded_log_height(5000, fixed_log2_rows=12)


def test_synthetic_events_are_well_formed():
    for chip in CHIPS:
        ev = tg.synthetic_events(chip, 2000, seed=5)
        opcode_word = 3 if chip in tg.FLOW_CHIPS else 2          # BranchEvent / JumpEvent vs AluEvent / MovCondEvent
        assert set(np.unique(ev[:, opcode_word]).tolist()) <= {tg.OPCODES[o] for o in tg.ALU_CHIPS[chip][1]}
        assert (ev[:, 0] < kb.P).all() and np.array_equal(ev[:, 1], ev[:, 0] + 4)
    ev = tg.synthetic_events("AddSub", 2000, seed=5)
    add = ev[:, 2] == 0
    assert np.array_equal(ev[add, 4], ev[add, 5] + ev[add, 6]) and np.array_equal(ev[~add, 4], ev[~add, 5] - ev[~add, 6])
    ev = tg.synthetic_events("Branch", 2000, seed=5)
    taken = ev[:, 2] != ev[:, 1] + 4
    assert 0.3 < taken.mean() < 0.7 and np.array_equal(ev[taken, 2], ev[taken, 1] + ev[taken, 6])
    for chip in tg.FLOW_CHIPS:          # every program counter is a KoalaBear word
        ev = tg.synthetic_events(chip, 2000, seed=5)
        assert (ev[:, :3] < kb.P).all()


def _alu_traces(oracle, n_add=1000, n_sll=300, seed=4):
    tr = {}
    for chip, n in (("AddSub", n_add), ("ShiftLeft", n_sll)):
        ev = tg.synthetic_events(chip, n, seed=seed)
        tr[chip] = (ev, oracle.alu_trace(chip, ev, 1 << tg.padded_log_height(n)))
    return tr


def test_generated_traces_satisfy_the_chips_real_constraints(oracle):
    """Rows identical to the reference's (events and padding rows) under the arithmetic constraints
    of AddSubChip::eval / ShiftLeft::eval restated in ziren_b200/synthetic.py: the restated prover
    and verifier accept them and reject single-cell corruptions."""
    from ziren_b200 import synthetic
    tr = {k: v[1] for k, v in _alu_traces(oracle).items()}
    case = synthetic.alu_case(tr, with_lookup_pair=False)
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    proof, _ = om.prove_shard(case.traces, case.public_values)
    ok, err = om.verify_shard(proof)
    assert ok, err
    for chip, row, col in (("AddSub", 5, 6), ("AddSub", 900, 2), ("ShiftLeft", 7, 31), ("ShiftLeft", 299, 30), ("ShiftLeft", 400, 22)):
        bad = {k: v.copy() for k, v in tr.items()}
        bad[chip][row, col] = (int(bad[chip][row, col]) + 1) % kb.P
        p2, _ = om.prove_shard({**case.traces, **bad}, case.public_values)
        assert not om.verify_shard(p2)[0], (chip, row, col)


def test_lt_traces_satisfy_the_restated_lt_constraints(oracle):
    """LtChip::eval's arithmetic constraints (sign handling, byte flags, inverse hint) over reference-identical rows."""
    from ziren_b200 import synthetic
    tr = {k: v[1] for k, v in _alu_traces(oracle, n_add=300, n_sll=100).items()}
    ev = tg.synthetic_events("Lt", 2000, seed=4)
    tr["Lt"] = oracle.alu_trace("Lt", ev, 1 << tg.padded_log_height(len(ev)))
    case = synthetic.alu_case(tr, with_lookup_pair=False)
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    proof, _ = om.prove_shard(case.traces, case.public_values)
    ok, err = om.verify_shard(proof)
    assert ok, err
    row = int(np.flatnonzero(ev[:, 5] != ev[:, 6])[-1])        # a row whose operands differ: every hint column is live
    for col in (4, 16, 22, 27, 29, 30):
        bad = {k: v.copy() for k, v in tr.items()}
        bad["Lt"][row, col] = (int(bad["Lt"][row, col]) + 1) % kb.P
        p2, _ = om.prove_shard(bad, case.public_values)
        assert not om.verify_shard(p2)[0], col


def test_all_alu_chips_traces_satisfy_their_restated_constraints(oracle):
    """The six ALU chips in one shard: reference-identical rows (events and padding rows) under the
    arithmetic constraints of each chip's Air::eval as restated in ziren_b200/synthetic.py."""
    from ziren_b200 import synthetic
    tr = {}
    for chip, n in (("AddSub", 300), ("ShiftLeft", 200), ("Lt", 700), ("ShiftRight", 1500), ("Bitwise", 100), ("CloClz", 700)):
        tr[chip] = oracle.alu_trace(chip, tg.synthetic_events(chip, n, seed=6), 1 << tg.padded_log_height(n))
    case = synthetic.alu_case(tr, with_lookup_pair=False)
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    proof, _ = om.prove_shard(case.traces, case.public_values)
    ok, err = om.verify_shard(proof)
    assert ok, err
    # event rows and padding rows (ShiftRight 1700, CloClz 900 are past the last event)
    for chip, row, col in (("ShiftRight", 900, 30), ("ShiftRight", 900, 22), ("ShiftRight", 1499, 46), ("ShiftRight", 1700, 10),
                           ("CloClz", 600, 10), ("CloClz", 900, 2), ("Bitwise", 50, 14)):
        bad = {k: v.copy() for k, v in tr.items()}
        bad[chip][row, col] = (int(bad[chip][row, col]) + 1) % kb.P
        p2, _ = om.prove_shard(bad, case.public_values)
        assert not om.verify_shard(p2)[0], (chip, row, col)


def test_branch_jump_movcond_traces_satisfy_their_restated_constraints(oracle):
    """BranchChip / JumpChip / MovCondChip arithmetic constraints (KoalaBear word range checkers, branch-taken
    logic, link value, per-byte is-zero hints) over reference-identical rows, including jump targets with the
    top byte 0x7f."""
    from ziren_b200 import synthetic
    tr = {}
    for chip, n in (("AddSub", 100), ("ShiftLeft", 100), ("Branch", 1500), ("Jump", 700), ("MovCond", 1500)):
        tr[chip] = oracle.alu_trace(chip, tg.synthetic_events(chip, n, seed=6), 1 << tg.padded_log_height(n))
    case = synthetic.alu_case(tr, with_lookup_pair=False)
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    proof, _ = om.prove_shard(case.traces, case.public_values)
    ok, err = om.verify_shard(proof)
    assert ok, err
    for chip, row, col in (("Branch", 900, 59), ("Branch", 900, 5), ("Branch", 900, 23), ("Branch", 1400, 60), ("Jump", 2, 19),
                           ("Jump", 600, 37), ("Jump", 650, 58), ("Jump", 900, 49), ("MovCond", 900, 18), ("MovCond", 901, 28),
                           ("MovCond", 1700, 29)):
        bad = {k: v.copy() for k, v in tr.items()}
        bad[chip][row, col] = (int(bad[chip][row, col]) + 1) % kb.P
        p2, _ = om.prove_shard(bad, case.public_values)
        assert not om.verify_shard(p2)[0], (chip, row, col)


# ---- GPU -------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from ziren_b200 import synthetic
    from ziren_b200.prover import B200Prover
    prover = B200Prover(synthetic.mini_case().machine, device=0)
    yield torch, prover
    prover.close()


def _gpu_trace(torch, prover, chip, ev, log_h, col_major, events_on_device=False):
    w, h = tg.width(chip), 1 << log_h
    out = torch.full((h * w,), -1, dtype=torch.int32, device="cuda")
    src = torch.from_numpy(ev.view(np.int32)).cuda() if events_on_device else ev
    prover.generate_alu_trace(chip, src, log_h, out, col_major=col_major)
    got = out.cpu().numpy().view(np.uint32)
    return got.reshape(w, h).T if col_major else got.reshape(h, w)


def _check_gpu_chip(gpu, oracle, chip):
    torch, prover = gpu
    gev, grows = np.array(GOLD[chip]["events"], np.uint32), np.array(GOLD[chip]["rows"], np.uint32)
    got = _gpu_trace(torch, prover, chip, gev, 7, col_major=False)
    assert np.array_equal(got[: len(gev)], grows)                     # the reference's own rows
    for n, log_h, cm, on_dev in [(5000, 13, False, False), (5000, 13, True, True), (129, 8, True, False), (256, 8, False, True),
                                 (1, 4, False, False), (0, 4, True, False), (3000, 12, False, True)]:
        ev = tg.synthetic_events(chip, n, seed=11 + n)
        want = kb.to_monty(oracle.alu_trace(chip, ev, 1 << log_h))
        assert np.array_equal(_gpu_trace(torch, prover, chip, ev, log_h, cm, on_dev), want), (chip, n, log_h, cm)


@pytest.mark.gpu
@pytest.mark.parametrize("chip", CHIPS[:6])
def test_gpu_trace_matches_oracle_and_golden(gpu, oracle, chip):
    _check_gpu_chip(gpu, oracle, chip)


@pytest.mark.gpu
def test_gpu_trace_large_and_layouts_agree(gpu, oracle):
    torch, prover = gpu
    ev = tg.synthetic_events("ShiftRight", (1 << 18) - 77, seed=2)
    rm = _gpu_trace(torch, prover, "ShiftRight", ev, 18, col_major=False)
    cm = _gpu_trace(torch, prover, "ShiftRight", ev, 18, col_major=True)
    assert np.array_equal(rm, cm)
    sample = np.r_[0:300, len(ev) - 300:len(ev)]
    assert np.array_equal(rm[sample], kb.to_monty(oracle.alu_trace("ShiftRight", ev[sample], len(sample))))
    assert np.array_equal(rm[len(ev):], kb.to_monty(oracle.alu_trace("ShiftRight", ev[:0], 77)))


@pytest.mark.gpu
def test_gpu_trace_errors(gpu):
    from ziren_b200.prover import ZkbError
    torch, prover = gpu
    out = torch.zeros(16 * 67, dtype=torch.int32, device="cuda")
    with pytest.raises(ZkbError, match="no row filler"):
        prover.generate_alu_trace("Byte", np.zeros((1, 7), np.uint32), 4, out)      # multiplicities, not rows of events (K7)
    with pytest.raises(ZkbError, match="more events than rows"):
        prover.generate_alu_trace("AddSub", tg.synthetic_events("AddSub", 17), 4, out)


@pytest.mark.gpu
@pytest.mark.parametrize("with_lookup_pair", [True, False])
def test_gpu_generated_traces_prove_bit_exact(gpu, oracle, with_lookup_pair):
    """events -> zkb200_generate_alu_trace (device-resident, row-major) -> zkb200_prove_shard: the proof
    equals the oracle's proof over the oracle's (= the reference's) rows and verifies."""
    from ziren_b200 import synthetic
    from ziren_b200.prover import B200Prover
    torch, _ = gpu
    tr = _alu_traces(oracle, n_add=3000, n_sll=700, seed=9)
    case = synthetic.alu_case({k: v[1] for k, v in tr.items()}, with_lookup_pair=with_lookup_pair)
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    want, _ = om.prove_shard(case.traces, case.public_values)
    prover = B200Prover(case.machine, device=0)
    try:
        pk = prover.setup({})
        dev_tr = {k: torch.from_numpy(kb.to_monty(v).view(np.int32)).cuda() for k, v in case.traces.items() if k not in tr}
        for chip, (ev, rows) in tr.items():
            log_h = tg.padded_log_height(len(ev))
            out = torch.empty((1 << log_h, tg.width(chip)), dtype=torch.int32, device="cuda")
            prover.generate_alu_trace(chip, ev, log_h, out)
            dev_tr[chip] = out
        got, _ = prover.prove_shard(pk, dev_tr, case.public_values)
        ok, err = om.verify_shard(got)
        assert ok, err
        assert np.array_equal(got, want)
        pk.free()
    finally:
        prover.close()


@pytest.mark.gpu
@pytest.mark.parametrize("chip", CHIPS[6:])
def test_gpu_trace_control_flow_chips(gpu, oracle, chip):
    """Branch, Jump, MovCond: kept last in the file, they were added after round 1's last GPU run."""
    _check_gpu_chip(gpu, oracle, chip)


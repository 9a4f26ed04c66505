"""Trace generation of the KeccakSponge precompile chip and the ZKB200_TRACE_EVENTS /
ZKB200_TRACE_COL_MAJOR inputs of zkb200_commit.

CPU tests: the oracle restatement (oracle/tracegen_keccak.h) against known answers (hashlib SHA3-256 through the
rows' own a''' columns), the memory columns against the REFERENCE'S OWN C++ (crates/core/machine/include/memory.hpp,
golden values in tests/golden/mem_access.json, live when oracle/_ref is present), the product's row filler compiled
for the host (ziren_b200/csrc/tracegen_keccak.cuh through tests/hostcheck) against the oracle.
GPU tests: the CUDA kernel through the C ABI against the oracle, bit-exact, both layouts; commit from event records."""
import ctypes
import hashlib
import json
import os

import numpy as np
import pytest

from ziren_b200 import field as kb
from ziren_b200 import keccak_sponge as ks
from ziren_b200 import tracegen as tg

HERE = os.path.dirname(os.path.abspath(__file__))
C = ks.Cols


def _host_rows(host, blocks, height):
    b = np.ascontiguousarray(blocks, dtype=np.uint32).reshape(-1, ks.REC_WORDS)
    out = np.zeros((height, ks.WIDTH), np.uint32)
    rc = host.hostcheck_keccak_rows(b.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(b)), ctypes.c_size_t(height),
                                    out.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    return out


def _state_after(row):
    """The 25 lanes a row hands to the next round: a_prime_prime_prime (columns.rs of keccak-air)."""
    limbs = [[int(row[C.A_PRIME_PRIME + 4 * i + l]) for l in range(4)] for i in range(25)]
    limbs[0] = [int(row[C.A_PRIME_PRIME_PRIME_0_0_LIMBS + l]) for l in range(4)]
    return [sum(v << (16 * k) for k, v in enumerate(l4)) for l4 in limbs]


def _sha3_256_through_rows(trace_fn, msg: bytes) -> bytes:
    rate = 136
    padded = bytearray(msg) + bytearray(rate - len(msg) % rate)
    padded[len(msg)] ^= 0x06
    padded[-1] ^= 0x80
    state = [0] * 25
    for off in range(0, len(padded), rate):
        block = np.frombuffer(bytes(padded[off:off + rate]), dtype="<u8")
        for i in range(17):
            state[i] ^= int(block[i])
        rec = np.zeros((1, ks.REC_WORDS), np.uint32)
        rec[0, ks.R_NBLOCKS] = 1
        for i in range(25):
            rec[0, ks.R_XORED_STATE + 2 * i] = state[i] & 0xFFFFFFFF
            rec[0, ks.R_XORED_STATE + 2 * i + 1] = state[i] >> 32
        rows = trace_fn(rec, 32)
        state = _state_after(rows[23])
    return b"".join(x.to_bytes(8, "little") for x in state[:4])


@pytest.mark.parametrize("msg", [b"", b"abc", b"zkb200 keccak sponge known answer", bytes(range(256)) * 2])
def test_oracle_permutation_rows_reproduce_sha3_256(oracle, msg):
    assert _sha3_256_through_rows(oracle.keccak_sponge_trace, msg) == hashlib.sha3_256(msg).digest()


def test_product_row_filler_reproduces_sha3_256(host):
    f = lambda rec, h: kb.from_monty(_host_rows(host, rec, h))
    msg = b"the product's own row filler, compiled for the host" * 7
    assert _sha3_256_through_rows(f, msg) == hashlib.sha3_256(msg).digest()


def test_numpy_keccak_f_known_answer():
    msg = b"event synthesis uses this permutation"
    blk = bytearray(136)
    blk[:len(msg)] = msg
    blk[len(msg)] ^= 0x06
    blk[135] ^= 0x80
    st = np.zeros((1, 25), np.uint64)
    st[0, :17] = np.frombuffer(bytes(blk), dtype="<u8")
    assert ks.keccak_f1600(st)[0, :4].tobytes() == hashlib.sha3_256(msg).digest()


def test_width_and_column_map(oracle, host):
    assert ks.WIDTH == oracle.KS_WIDTH == host.hostcheck_keccak_width() == 3531
    # KeccakCols: 24 + 1 + 100 + 100 + 320 + 320 + 1600 + 100 + 64 + 4; sponge: 36*9 + 13 + 50*4 + 36*4 + 9 + 16*13
    assert C.NUM_KECCAK_COLS == 24 + 1 + 100 + 100 + 320 + 320 + 1600 + 100 + 64 + 4
    assert C.WIDTH - C.NUM_KECCAK_COLS == 36 * 9 + 13 + 50 * 4 + 36 * 4 + 9 + 16 * 13
    assert ks.REC_WORDS * 4 == 1536 and ks.R_WRITES + 6 * 16 <= ks.REC_WORDS


def test_memory_columns_match_the_reference_cpp(oracle):
    """MemoryAccessCols::populate_access against the reference's own memory.hpp: golden values (generated here by
    tests/golden/gen_mem_golden.py from oracle/_ref) and, when oracle/_ref is present, live."""
    gold = json.load(open(os.path.join(HERE, "golden", "mem_access.json")))
    for rec, cols in zip(gold["records"], gold["read_cols"]):
        assert np.array_equal(kb.to_monty(oracle.mem_access(*rec)), np.array(cols, np.uint32))
    if oracle.ref_mem_access(1, 1, 2, 1, 1) is None:
        return
    rng = np.random.default_rng(7)
    for _ in range(3000):
        sh = int(rng.integers(1, 1 << 16))
        same = bool(rng.integers(0, 2))
        psh = sh if same else int(rng.integers(0, sh))
        ts = int(rng.integers(1, 1 << 24))
        pts = int(rng.integers(0, ts)) if same else int(rng.integers(0, 1 << 24))
        v = int(rng.integers(0, 1 << 32))
        assert np.array_equal(kb.to_monty(oracle.mem_access(v, sh, ts, psh, pts)), oracle.ref_mem_access(v, sh, ts, psh, pts))


@pytest.mark.parametrize("per,shard,height", [([1], 1, 32), ([1, 2, 3, 1, 4, 1, 2], 3, 512), ([5] * 9, 2, 2048), ([], 1, 64)])
def test_product_row_filler_matches_oracle(oracle, host, per, shard, height):
    b = ks.synthetic_blocks(len(per), per if per else 1, seed=len(per), shard=shard)
    want = kb.to_monty(oracle.keccak_sponge_trace(b, height))
    assert np.array_equal(_host_rows(host, b, height), want)


def test_rows_have_the_structure_event_to_rows_gives_them(oracle):
    per = [1, 3, 2]
    b = ks.synthetic_blocks(3, per, seed=5, shard=4)
    t = oracle.keccak_sponge_trace(b, 256)
    n_real = 24 * sum(per)
    rounds = np.arange(256) % 24
    assert np.array_equal(np.argmax(t[:, C.STEP_FLAGS:C.STEP_FLAGS + 24], axis=1), rounds)       # also on the dummy rows
    assert (t[:, C.STEP_FLAGS:C.STEP_FLAGS + 24].sum(axis=1) == 1).all() and (t[:, C.EXPORT] == 0).all()
    assert (t[:n_real, C.IS_REAL] == 1).all() and (t[n_real:, C.NUM_KECCAK_COLS:] == 0).all()
    assert np.array_equal(t[:n_real, C.READ_BLOCK], (rounds[:n_real] == 0).astype(np.uint32))
    blk = np.repeat(b[:, ks.R_BLOCK], 24)
    nb = np.repeat(b[:, ks.R_NBLOCKS], 24)
    r = rounds[:n_real]
    assert np.array_equal(t[:n_real, C.RECEIVE_SYSCALL], ((blk == 0) & (r == 0)).astype(np.uint32))
    assert np.array_equal(t[:n_real, C.WRITE_OUTPUT], ((blk == nb - 1) & (r == 23)).astype(np.uint32))
    assert np.array_equal(t[:n_real, C.IS_ABSORBED], ((blk != nb - 1) & (r == 23)).astype(np.uint32))
    assert np.array_equal(t[:n_real, C.ALREADY_ABSORBED_U32S], blk * 36)
    # the dummy rows are the zero-input permutation's rows by row index mod 24
    zero = oracle.keccak_sponge_trace(np.zeros((0, ks.REC_WORDS), np.uint32), 32)
    assert np.array_equal(t[n_real:, :C.NUM_KECCAK_COLS], zero[rounds[n_real:], :C.NUM_KECCAK_COLS])
    # a row's `a` is the previous round's a''' and, across the blocks of one event, the next block's original_state
    # is this block's output (what air.rs:230-262 constrains under is_absorbed)
    for row in (5, 23 + 24, 24 * 2 + 23):
        limbs = np.array(_state_after(t[row]), dtype=np.uint64)
        if rounds[row] != 23:
            nxt = [sum(int(t[row + 1, C.A + 4 * i + l]) << (16 * l) for l in range(4)) for i in range(25)]
            assert limbs.tolist() == nxt
        else:
            assert t[row, C.IS_ABSORBED] == 1
            words = t[row + 1, C.ORIGINAL_STATE:C.ORIGINAL_STATE + 200].reshape(50, 4)
            got = [int(w[0]) | int(w[1]) << 8 | int(w[2]) << 16 | int(w[3]) << 24 for w in words]
            assert [int(x) & 0xFFFFFFFF for x in limbs.tolist() for _ in (0,)] == got[0::2]
            assert [int(x) >> 32 for x in limbs.tolist()] == got[1::2]
    # the last row of every event writes the output the sponge ends with
    last = 24 * 1 - 1
    out_words = t[last, C.OUTPUT_MEM:C.OUTPUT_MEM + 16 * 13].reshape(16, 13)[:, 4:8]
    out_vals = [int(w[0]) | int(w[1]) << 8 | int(w[2]) << 16 | int(w[3]) << 24 for w in out_words]
    st = _state_after(t[last])
    assert out_vals == [(st[i // 2] >> (32 * (i % 2))) & 0xFFFFFFFF for i in range(16)]


def _real_case(oracle, per=(1, 2, 1, 1), log_cpu=10, seed=3):
    from ziren_b200 import synthetic
    b = ks.synthetic_blocks(len(per), list(per), seed=seed, shard=1)
    t = oracle.keccak_sponge_trace(b, 1 << ks.padded_log_height(len(b)))
    return b, synthetic.keccak_real_case(b, t, log_cpu=log_cpu, num_queries=6, pow_bits=3)


def test_real_keccak_sponge_air_accepts_the_generated_rows(oracle):
    """KeccakSpongeChip::eval restated (ziren_b200/keccak_air.py: the p3 KeccakAir constraints, the sponge rules, 357
    lookups balanced by Byte / MemoryLocalPrecompile / SyscallPrecompile tables) over the row filler's rows: proved and
    verified on the CPU, single-cell corruptions rejected.  The chip's cost equals the reference's cost table."""
    from ziren_b200 import keccak_air
    chip = keccak_air.keccak_sponge_chip()
    # crates/core/executor/src/artifacts/mips_costs.json: "KeccakSponge": 102216 per 24 rows
    assert chip.cost * 24 == 102216 and chip.main_width == 3531 and chip.perm_width_ef == 180 and chip.log_quotient_degree == 1
    assert len(chip.builder.sends) + len(chip.builder.receives) == 357
    b, case = _real_case(oracle)
    om = oracle.OracleMachine(case.machine)
    om.setup(case.prep)
    proof, _ = om.prove_shard(case.traces, case.public_values)
    ok, err = om.verify_shard(proof)
    assert ok, err
    t = case.traces["KeccakSponge"]
    cells = [(0, C.A_PRIME + 7), (5, C.C + 64), (23, C.A_PRIME_PRIME_PRIME_0_0_LIMBS), (24, C.ORIGINAL_STATE + 3), (47, C.IS_ABSORBED),
             (0, C.BLOCK_MEM + C.M_DIFF16), (0, C.XORED_GENERAL_RATE + 5), (71, C.OUTPUT_MEM + 4), (3, C.CLK), (100, C.STEP_FLAGS + 4),
             (126, C.A_PRIME_PRIME + 9)]
    for row, col in cells:
        bad = dict(case.traces)
        bad["KeccakSponge"] = t.copy()
        bad["KeccakSponge"][row, col] = (int(t[row, col]) + 1) % kb.P
        try:
            p2, _ = om.prove_shard(bad, case.public_values)
        except RuntimeError:
            continue                      # the oracle prover refuses (final polynomial not constant)
        assert not om.verify_shard(p2)[0], (row, col)
    # a missing byte-table multiplicity unbalances the shard's local cumulative sums
    bad = dict(case.traces)
    bad["Byte"] = case.traces["Byte"].copy()
    i = int(np.flatnonzero(bad["Byte"][:, 1])[0])
    bad["Byte"][i, 1] -= 1
    p2, _ = om.prove_shard(bad, case.public_values)
    assert not om.verify_shard(p2)[0]


def test_padded_height_rule():
    assert [ks.padded_log_height(n) for n in (0, 1, 2, 5, 6, 10922, 10923)] == [0, 5, 6, 7, 8, 18, 19]
    assert ks.padded_log_height(3, fixed_log2_rows=10) == 10
    with pytest.raises(ValueError):
        ks.padded_log_height(100, fixed_log2_rows=10)


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


def _gpu_trace(torch, prover, blocks, log_h, col_major, on_device=False):
    h = 1 << log_h
    out = torch.full((h * ks.WIDTH,), -1, dtype=torch.int32, device="cuda")
    src = torch.from_numpy(blocks.view(np.int32)).cuda() if on_device and len(blocks) else blocks
    prover.generate_keccak_sponge_trace(src, log_h, out, col_major=col_major)
    got = out.cpu().numpy().view(np.uint32)
    return got.reshape(ks.WIDTH, h).T if col_major else got.reshape(h, ks.WIDTH)


@pytest.mark.gpu
@pytest.mark.parametrize("per,log_h,col_major,on_device", [([1], 5, True, False), ([1, 2, 3, 1, 4, 1, 2], 9, True, True),
                                                           ([5] * 9, 11, False, False), ([], 6, True, False),
                                                           ([2] * 170, 13, True, True), ([1] * 5, 7, False, True)])
def test_gpu_keccak_trace_matches_oracle(gpu, oracle, per, log_h, col_major, on_device):
    torch, prover = gpu
    b = ks.synthetic_blocks(len(per), per if per else 1, seed=40 + len(per), shard=2)
    want = kb.to_monty(oracle.keccak_sponge_trace(b, 1 << log_h))
    assert np.array_equal(_gpu_trace(torch, prover, b, log_h, col_major, on_device), want)


@pytest.mark.gpu
def test_gpu_keccak_trace_reproduces_sha3_256(gpu):
    torch, prover = gpu
    f = lambda rec, h: kb.from_monty(_gpu_trace(torch, prover, rec, 5, True))
    msg = b"through the CUDA kernel" * 20
    assert _sha3_256_through_rows(f, msg) == hashlib.sha3_256(msg).digest()


@pytest.mark.gpu
def test_gpu_keccak_trace_errors(gpu):
    from ziren_b200.prover import ZkbError
    torch, prover = gpu
    out = torch.zeros(32 * ks.WIDTH, dtype=torch.int32, device="cuda")
    with pytest.raises(ZkbError, match="more rows than"):
        prover.generate_keccak_sponge_trace(ks.synthetic_blocks(2, 1), 5, out)


@pytest.mark.gpu
def test_commit_from_events_errors(gpu, oracle):
    """ZKB200_TRACE_EVENTS for a chip without a row filler, too many events, a wrong width: MachineProver::Error, no crash."""
    from ziren_b200 import synthetic
    from ziren_b200.prover import B200Prover, EventTrace, ZkbError
    case = synthetic.mini_case()
    prover = B200Prover(case.machine, device=0)
    try:
        inputs = {k: kb.to_monty(v) for k, v in case.traces.items()}
        bad = dict(inputs)
        bad["Byte"] = EventTrace(np.zeros((4, 7), np.uint32), 16, case.traces["Byte"].shape[1])     # a table without a row filler
        with pytest.raises(ZkbError, match="no row filler"):
            prover.commit(bad, case.public_values)
        bad = dict(inputs)
        bad["AddSub"] = EventTrace(tg.synthetic_events("AddSub", 40), 5, case.traces["AddSub"].shape[1])
        with pytest.raises(ZkbError, match="another width|more event rows"):
            prover.commit(bad, case.public_values)
        data = prover.commit(inputs, case.public_values)      # the context is still usable
        data.free()
    finally:
        prover.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["rows_host", "events_host", "events_device"])
def test_real_keccak_shard_proves_bit_exact(gpu, oracle, mode):
    """A shard with the REAL KeccakSponge chip: uploaded rows, or event records handed to zkb200_commit (the table is
    generated inside the commit) - the proof is the oracle's proof over the oracle's rows, word for word."""
    from ziren_b200.prover import B200Prover, EventTrace
    torch, _ = gpu
    b, case = _real_case(oracle, per=(1, 3, 2, 4, 1), log_cpu=11, seed=8)
    om = oracle.OracleMachine(case.machine)
    om.setup(case.prep)
    want, _ = om.prove_shard(case.traces, case.public_values)
    prover = B200Prover(case.machine, device=0)
    try:
        pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
        inputs = {k: kb.to_monty(v) for k, v in case.traces.items()}
        if mode != "rows_host":
            ev = torch.from_numpy(b.view(np.int32)).cuda() if mode == "events_device" else b
            inputs["KeccakSponge"] = EventTrace(ev, ks.padded_log_height(len(b)), ks.WIDTH)
        got, _ = prover.prove_shard(pk, inputs, case.public_values)
        ok, err = om.verify_shard(got)
        assert ok, err
        assert np.array_equal(got, want)
        pk.free()
    finally:
        prover.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["events_host", "events_device", "col_major"])
def test_commit_from_events_proves_bit_exact(gpu, oracle, mode):
    """zkb200_commit fed with EVENT RECORDS (ZKB200_TRACE_EVENTS: the row fillers run inside the commit, nothing is
    transposed) or with device column-major tables (ZKB200_TRACE_COL_MAJOR): the proof equals the oracle's proof over
    the reference-identical rows."""
    from ziren_b200 import synthetic
    from ziren_b200.prover import B200Prover, ColMajorTrace, EventTrace
    torch, _ = gpu
    tr = {}
    for chip, n in (("AddSub", 3000), ("ShiftLeft", 700), ("Lt", 1200)):
        ev = tg.synthetic_events(chip, n, seed=9)
        tr[chip] = (ev, oracle.alu_trace(chip, ev, 1 << tg.padded_log_height(n)))
    case = synthetic.alu_case({k: v[1] for k, v in tr.items()}, with_lookup_pair=True)
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    want, _ = om.prove_shard(case.traces, case.public_values)
    prover = B200Prover(case.machine, device=0)
    try:
        pk = prover.setup({})
        inputs = {k: kb.to_monty(v) for k, v in case.traces.items() if k not in tr}
        keep = []
        for chip, (ev, rows) in tr.items():
            log_h = tg.padded_log_height(len(ev))
            if mode == "events_host":
                inputs[chip] = EventTrace(ev, log_h, tg.width(chip))
            elif mode == "events_device":
                d = torch.from_numpy(ev.view(np.int32)).cuda()
                keep.append(d)
                inputs[chip] = EventTrace(d, log_h, tg.width(chip))
            else:
                out = torch.empty((tg.width(chip), 1 << log_h), dtype=torch.int32, device="cuda")
                prover.generate_alu_trace(chip, ev, log_h, out, col_major=True)
                inputs[chip] = ColMajorTrace(out, 1 << log_h, tg.width(chip))
        got, _ = prover.prove_shard(pk, inputs, case.public_values)
        ok, err = om.verify_shard(got)
        assert ok, err
        assert np.array_equal(got, want)
        pk.free()
    finally:
        prover.close()

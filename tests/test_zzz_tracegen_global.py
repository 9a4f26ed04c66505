"""Trace generation of the Global chip: every global lookup lifted to a point of the septic curve,
the points summed along the table, 99 columns.

CPU tests: the F_p^7 / curve primitives of the oracle (oracle/tracegen_global.h, the reference's formulation: Cipolla square
root, left-to-right sum) and of the product compiled for the host (ziren_b200/csrc/tracegen_global.cuh: Tonelli-Shanks, chunked
scan) against golden values written by the REFERENCE'S OWN C++ (crates/core/machine/include/kb31_septic_extension_t.hpp;
tests/golden/septic.json), against that C++ live when oracle/_ref is present, and against the properties the reference's own
tests assert (crates/stark/src/septic_extension.rs tests, septic_curve.rs tests); the rows of the product, walked in the
kernels' order, against the oracle's; the rows against the chip's constraints.  GPU: the CUDA kernels through the C ABI."""
import ctypes
import json
import os

import numpy as np
import pytest

from ziren_b200 import field as kb
from ziren_b200 import tracegen as tg

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "septic.json")))
P = kb.P
START = [637514027, 1595065213, 1998064738, 72333738, 1211544370, 822986770, 1518535784,
         1604177449, 90440090, 259343427, 140470264, 1162099742, 941559812, 1064053343]       # septic_digest.rs:9-14
DUMMY = [1706420302, 1319108093, 148224806, 26874985, 1766171812, 1645633948, 2028659224,
         942390502, 1239997438, 458866455, 1843332012, 1309764648, 572807436, 74267719]       # septic_curve.rs:14-19


# F_p^7 = F_p[z] / (z^7 + 2z - 8) in plain Python integers: the independent check of both implementations
def s_mul(a, b):
    r = [0] * 13
    for i in range(7):
        for j in range(7):
            r[i + j] += int(a[i]) * int(b[j])
    out = r[:7]
    for i in range(7, 13):
        out[i - 7] += 8 * r[i]
        out[i - 6] -= 2 * r[i]
    return [x % P for x in out]


def s_add(a, b):
    return [(int(x) + int(y)) % P for x, y in zip(a, b)]


def s_sub(a, b):
    return [(int(x) - int(y)) % P for x, y in zip(a, b)]


def curve_rhs(x):
    return s_sub(s_add(s_mul(s_mul(x, x), x), s_mul(x, [0, 3, 0, 0, 0, 0, 0])), [3, 0, 0, 0, 0, 0, 0])


def on_curve(pt):
    return s_mul(pt[7:], pt[7:]) == curve_rhs(pt[:7])


def host_septic(host, op, a, b=None):
    from oracle import oracle_ffi
    return oracle_ffi._septic_call(host.hostcheck_septic_op, op, a, b)


def impls(oracle, host):
    return (("oracle", oracle.septic_op), ("product", lambda op, a, b=None: host_septic(host, op, a, b)))


def test_primitives_match_reference_golden(oracle, host):
    for name, f in impls(oracle, host):
        for c in GOLD["cases"]:
            a, b = np.array(c["a"], np.uint32), np.array(c["b"], np.uint32)
            assert f("mul", a, b).tolist() == c["mul"] == s_mul(a, b), name
            assert f("inv", a).tolist() == c["inv"], name
            assert s_mul(a, c["inv"]) == [1, 0, 0, 0, 0, 0, 0]
            assert f("frobenius", a).tolist() == c["frobenius"], name
            assert f("double_frobenius", a).tolist() == c["double_frobenius"], name
            assert f("curve_formula", a).tolist() == c["curve_formula"] == curve_rhs(a), name
            # a square root is determined up to its sign (lift_x normalises it afterwards)
            root = f("sqrt", np.array(s_mul(a, a), np.uint32))
            assert root is not None and (root.tolist() == c["sqrt_of_square"] or s_add(root, c["sqrt_of_square"]) == [0] * 7), name
            assert s_mul(root, root) == s_mul(a, a)
            assert (f("sqrt", a) is not None) == c["a_is_square"], name
        for c in GOLD["curve_add"]:
            assert on_curve(c["p"]) and on_curve(c["q"]) and on_curve(c["sum"])
            assert f("curve_add", np.array(c["p"], np.uint32), np.array(c["q"], np.uint32)).tolist() == c["sum"], name


def test_primitives_match_reference_cpp_live(oracle, host):
    rng = np.random.default_rng(3)
    try:
        oracle.ref_septic_op("mul", np.ones(7, np.uint32), np.ones(7, np.uint32))
    except LookupError:
        pytest.skip("oracle/_ref not built (no reference checkout on this machine)")
    for _ in range(200):
        a, b = rng.integers(0, P, 7).astype(np.uint32), rng.integers(0, P, 7).astype(np.uint32)
        for name, f in impls(oracle, host):
            for op in ("mul", "inv", "frobenius", "double_frobenius", "curve_formula"):
                assert np.array_equal(f(op, a, b if op == "mul" else None), oracle.ref_septic_op(op, a, b if op == "mul" else None)), (name, op)
            want = oracle.ref_septic_op("sqrt", a)
            got = f("sqrt", a)
            assert (got is None) == (want is None), name
            if got is not None:
                assert np.array_equal(got, want) or s_add(got, want) == [0] * 7, name


def test_properties_the_reference_tests_assert(oracle, host):
    """septic_extension.rs tests test_inv / test_legendre / test_sqrt, septic_curve.rs test_lift_x / test_double, and the two
    constant points lying on the curve (septic_digest.rs test_const_points)."""
    assert on_curve(START) and on_curve(DUMMY)
    for name, f in impls(oracle, host):
        g, b = [2, 1, 0, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0, 0]            # SepticExtension::GENERATOR
        for i in range(1, 64):
            b = s_mul(b, g)
            root = f("sqrt", np.array(b, np.uint32))
            assert (root is not None) == (i % 2 == 0), name
            if root is not None:
                assert s_mul(root, root) == b
        # frobenius is a ring homomorphism fixing the base field, and applying it seven times is the identity
        a = np.array([5, 6, 17, 91, 37, 35, 33], np.uint32)
        x = a
        for _ in range(7):
            x = f("frobenius", x)
        assert np.array_equal(x, a), name
        assert np.array_equal(f("double_frobenius", a), f("frobenius", f("frobenius", a))), name
        # doubling: P + P is on the curve and equals (P + Q) + (P - Q) for another point Q
        p, q = np.array(GOLD["curve_add"][0]["p"], np.uint32), np.array(GOLD["curve_add"][0]["q"], np.uint32)
        dbl = f("curve_add", p, p)
        assert on_curve(dbl.tolist()), name
        neg_q = np.array(list(q[:7]) + [(P - int(v)) % P for v in q[7:]], np.uint32)
        assert np.array_equal(f("curve_add", f("curve_add", p, q), f("curve_add", p, neg_q)), dbl), name
        # P + (-P) is the point at infinity (stored as zeros), infinity is the neutral element
        neg_p = np.array(list(p[:7]) + [(P - int(v)) % P for v in p[7:]], np.uint32)
        inf = f("curve_add", p, neg_p)
        assert not inf.any() and np.array_equal(f("curve_add", inf, q), q) and np.array_equal(f("curve_add", q, inf), q), name


def _host_rows(host, ev, height, col_major=False):
    ev = np.ascontiguousarray(ev, dtype=np.uint32).reshape(-1, tg.GLOBAL_EVENT_WORDS)
    out = np.full(height * tg.GLOBAL_WIDTH, 0xFFFFFFFF, np.uint32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert host.hostcheck_global_rows(p(ev), ctypes.c_size_t(len(ev)), ctypes.c_size_t(height), p(out), int(col_major)) == 0
    return out.reshape(tg.GLOBAL_WIDTH, height).T if col_major else out.reshape(height, tg.GLOBAL_WIDTH)


@pytest.mark.parametrize("n,height", [(0, 16), (1, 16), (31, 32), (32, 32), (33, 64), (1023, 1024), (1024, 1024), (1025, 2048), (3000, 4096)])
def test_product_rows_match_oracle(oracle, host, n, height):
    """Chunk boundaries of the scan (32 points per chunk, the start point shifts everything by one): one, two and three levels."""
    assert tg.width("Global") == oracle.GLOBAL_WIDTH == host.hostcheck_global_width() == 99
    ev = tg.synthetic_global_events(n, seed=60 + n)
    want = kb.to_monty(oracle.global_trace(ev, height))
    assert np.array_equal(_host_rows(host, ev, height), want)
    if n in (33, 1025):
        assert np.array_equal(_host_rows(host, ev, height, col_major=True), want)


def test_rows_hold_the_chip_constraints(oracle):
    """GlobalLookupOperation::eval_single_digest and GlobalAccumulationOperation::eval_accumulation, in integers."""
    n, h = 200, 256
    ev = tg.synthetic_global_events(n, seed=9)
    t = oracle.global_trace(ev, h)
    is_receive, kind = ev[:, 7] & 0xFF, (ev[:, 7] >> 8) & 0xFF
    assert np.array_equal(t[:n, :7], ev[:, :7] % P) and np.array_equal(t[:n, 7], kind)
    assert np.array_equal(t[:n, 61], is_receive) and np.array_equal(t[:n, 62], 1 - is_receive) and (t[:n, 63] == 1).all()
    offset = sum(t[:n, 8 + k].astype(np.uint32) << k for k in range(8))
    assert (offset > 0).any()
    y6v = sum(t[:n, 30 + k].astype(np.uint64) << np.uint64(k) for k in range(30))
    prev = START
    for i in range(n):
        x, y = t[i, 16:23].tolist(), t[i, 23:30].tolist()
        # the map to the curve: x = (message[0] + kind * 2^16, message[1..5], message[6] * 256 + offset), (x, y) on the curve
        assert x[0] == (int(ev[i, 0]) + (int(kind[i]) << 16)) % P and x[1:6] == (ev[i, 1:6] % P).tolist()
        assert x[6] == (int(ev[i, 6]) * 256 + int(offset[i])) % P
        assert on_curve(x + y)
        # the sign of y: a receive has 1 <= y6 <= (p - 1) / 2, a send (p + 1) / 2 <= y6 <= p - 1
        assert y[6] == (1 if is_receive[i] else (P + 1) // 2) + int(y6v[i]) and int(y6v[i]) < (P - 1) // 2
        top = int(t[i, 53:60].sum())
        assert int(t[i, 60]) * ((top - 7) % P) % P == 1
        # the accumulation: initial digest = the previous row's sum, sum = initial + point (checked through the sum checkers)
        init, cum = t[i, 64:78].tolist(), t[i, 85:99].tolist()
        assert init == prev and (t[i, 78:85] == 0).all() and on_curve(cum)
        dx, dy = s_sub(x, init[:7]), s_sub(y, init[7:])
        assert s_sub(s_mul(s_add(s_add(init[:7], x), cum[:7]), s_mul(dx, dx)), s_mul(dy, dy)) == [0] * 7            # sum_checker_x
        assert s_sub(s_mul(s_add(init[7:], cum[7:]), dx), s_mul(dy, s_sub(init[:7], cum[:7]))) == [0] * 7           # sum_checker_y
        prev = cum
    # dummy rows: the dummy point, the final digest carried along, the witness of the "sum" with the dummy point
    fin = prev
    assert (t[n:, :16] == 0).all() and (t[n:, 30:64] == 0).all() and (t[n:, 16:30] == np.array(DUMMY, np.uint32)).all()
    assert (t[n:, 64:78] == np.array(fin, np.uint32)).all() and (t[n:, 85:99] == np.array(fin, np.uint32)).all()
    dx, dy = s_sub(DUMMY[:7], fin[:7]), s_sub(DUMMY[7:], fin[7:])
    chk = s_sub(s_mul(s_add(s_add(fin[:7], DUMMY[:7]), fin[:7]), s_mul(dx, dx)), s_mul(dy, dy))
    assert (t[n:, 78:85] == np.array(chk, np.uint32)).all()
    # no event at all: the reference's scan is empty and the final digest is the dummy point
    t0 = oracle.global_trace(ev[:0], 16)
    assert (t0[:, 64:78] == np.array(DUMMY, np.uint32)).all() and (t0[:, 78:85] == 0).all()


def test_the_sum_does_not_depend_on_the_scan_shape(oracle, host):
    """The digest of a table is the digest of its halves chained: the last cumulative sum of the whole table, computed by the
    product's chunked scan, against a left-to-right sum of the oracle's points through the oracle's addition."""
    ev = tg.synthetic_global_events(700, seed=12)
    rows = kb.from_monty(_host_rows(host, ev, 1024))
    acc = np.array(START, np.uint32)
    for i in range(700):
        acc = oracle.septic_op("curve_add", acc, rows[i, 16:30].copy())
        if i in (0, 31, 32, 63, 64, 699):
            assert np.array_equal(acc, rows[i, 85:99])
    assert np.array_equal(acc, rows[1023, 64:78])


def test_generated_rows_satisfy_the_restated_air(oracle):
    """Rows from trace generation under GlobalChip::eval restated as data (ziren_b200/synthetic.py _global_chip): the restated
    prover and verifier accept them - the shard's global cumulative sum is the table's final digest - and reject single-cell
    corruptions of every column group."""
    from ziren_b200 import synthetic
    ev = tg.synthetic_global_events(100, seed=2)
    rows = oracle.global_trace(ev, 128)
    case = synthetic.global_case(rows)
    chip = case.machine.chip("Global")
    assert chip.log_quotient_degree == 1 and chip.global_scope and chip.num_constraints == 134
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    proof, _ = om.prove_shard(case.traces, case.public_values)
    ok, err = om.verify_shard(proof)
    assert ok, err
    # message, offset bit, x, y, y6 bit, witness, is_real, initial digest (first row and later), checker, cumulative sum, dummy rows
    for row, col in ((5, 2), (9, 9), (5, 16), (5, 23), (11, 40), (3, 60), (99, 63), (0, 64), (50, 71), (7, 80), (7, 90), (100, 64), (127, 85)):
        bad = rows.copy()
        bad[row, col] = (int(bad[row, col]) + 1) % P
        p2, _ = om.prove_shard({**case.traces, "Global": bad}, case.public_values)
        assert not om.verify_shard(p2)[0], (row, col)


def _memory_shard(oracle, first_shard, n_init=40, n_fin=70, log_init=6, log_fin=7, seed=1, n_syscalls=0):
    """Memory initialisation / finalisation events of one shard, the global lookups they emit
    (MemoryGlobalChip::generate_dependencies) and the three tables from trace generation."""
    from ziren_b200 import synthetic
    init = tg.synthetic_memory_global_events(n_init, seed=seed)
    init[:, 2], init[:, 3] = 0, 1                        # MemoryGlobalInit constrains timestamp = 1
    fin = tg.synthetic_memory_global_events(n_fin, seed=seed + 1)
    if first_shard:                                      # previous address 0: the table starts with register 0 = 0
        init[0, :2] = 0
        fin[0, :2] = 0
        prev_init = prev_fin = 0
    else:
        prev_init, prev_fin = int(init[:, 0].min()) - 1, int(fin[:, 0].min()) - 3
    rec_init, rec_fin = tg.memory_global_records(init, prev_init), tg.memory_global_records(fin, prev_fin)
    gev = np.concatenate([tg.memory_global_lookup_events(init, False), tg.memory_global_lookup_events(fin, True)])
    sys_ev = tg.synthetic_syscall_events(n_syscalls, seed=seed, kind="core")
    if n_syscalls:                                       # a core shard's syscall table: two more global lookups per row
        gev = np.concatenate([gev, tg.syscall_global_lookup_events(sys_ev)])
    log_g = tg.padded_log_height(len(gev))
    rows = {"MemoryGlobalInit": oracle.memory_global_trace(rec_init[:, :4], prev_init, 1 << log_init),
            "MemoryGlobalFinalize": oracle.memory_global_trace(rec_fin[:, :4], prev_fin, 1 << log_fin),
            "Global": oracle.global_trace(gev, 1 << log_g)}
    events = {"MemoryGlobalInit": (rec_init, log_init), "MemoryGlobalFinalize": (rec_fin, log_fin), "Global": (gev, log_g)}
    if n_syscalls:
        log_s = tg.padded_log_height(n_syscalls)
        rows["SyscallCore"] = oracle.chip_trace("SyscallCore", sys_ev, 1 << log_s)
        events["SyscallCore"] = (sys_ev, log_s)
    case = synthetic.memory_global_case(rows["MemoryGlobalInit"], rows["MemoryGlobalFinalize"], rows["Global"], prev_init, prev_fin,
                                        syscall_rows=rows.get("SyscallCore"))
    return case, rows, events


@pytest.mark.parametrize("first_shard", [False, True])
def test_memory_tables_and_global_table_satisfy_their_airs_and_the_lookup_between_them(oracle, first_shard):
    """MemoryGlobalInit, MemoryGlobalFinalize, Global (and, in the later shard, SyscallCore with its two global lookups per row:
    SyscallChip::eval restated) from trace generation under MemoryGlobalChip::eval / GlobalChip::eval
    restated as data, tied by the real lookup: every real memory row sends (shard, timestamp, addr, value bytes, is_send,
    is_receive, Memory), Global receives its messages - so the Global table's events must be exactly the memory tables'
    rows.  The restated prover and verifier accept the shard and reject single-cell corruptions (a flipped value bit breaks
    the lookup balance, the others a constraint)."""
    case, rows, _ = _memory_shard(oracle, first_shard, n_syscalls=0 if first_shard else 25)
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    proof, _ = om.prove_shard(case.traces, case.public_values)
    ok, err = om.verify_shard(proof)
    assert ok, err
    # addr, address bit, lt flag, value bit, is_next_comp, is_first_comp, is_last_addr; a Global message word
    for name, row, col in (("MemoryGlobalInit", 3, 2), ("MemoryGlobalInit", 7, 40), ("MemoryGlobalFinalize", 9, 10), ("MemoryGlobalFinalize", 5, 80),
                           ("MemoryGlobalFinalize", 4, 106), ("MemoryGlobalFinalize", 0, 109), ("MemoryGlobalFinalize", 69, 110), ("Global", 4, 3),
                           ("SyscallCore", 3, 4), ("SyscallCore", 6, 7), ("SyscallCore", 30, 9)):       # an argument half-word, a result half-word, is_linux of a padding row
        if name not in rows:
            continue
        bad = rows[name].copy()
        bad[row, col] = (int(bad[row, col]) + 1) % P
        p2, _ = om.prove_shard({**case.traces, name: bad}, case.public_values)
        assert not om.verify_shard(p2)[0], (name, row, col)


def _precompile_shard(oracle, n, seed=3):
    from ziren_b200 import synthetic
    ev = tg.synthetic_syscall_events(n, seed=seed, kind="precompile")
    gev = tg.syscall_global_lookup_events(ev, precompile=True)
    log_s, log_g = tg.padded_log_height(n), tg.padded_log_height(len(gev))
    rows = {"SyscallPrecompile": oracle.chip_trace("SyscallPrecompile", ev, 1 << log_s), "Global": oracle.global_trace(gev, 1 << log_g)}
    case = synthetic.syscall_precompile_case(rows["SyscallPrecompile"], rows["Global"])
    return case, rows, {"SyscallPrecompile": (ev, log_s), "Global": (gev, log_g)}


def test_precompile_shard_tables_satisfy_their_airs_and_the_lookup_between_them(oracle):
    """SyscallPrecompile (one SyscallEvent per precompile event, linux results carried in a_record) and the Global table of the
    lookups it emits - received here, sent by the core shard - under the restated constraints and the real lookup."""
    case, rows, _ = _precompile_shard(oracle, 40)
    assert (rows["SyscallPrecompile"][:40, 9] == 1).any() and (rows["Global"][:80, 61] == 1).all()
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    proof, _ = om.prove_shard(case.traces, case.public_values)
    ok, err = om.verify_shard(proof)
    assert ok, err
    for name, row, col in (("SyscallPrecompile", 3, 5), ("SyscallPrecompile", 2, 2), ("Global", 7, 61), ("Global", 9, 7)):
        bad = rows[name].copy()
        bad[row, col] = (int(bad[row, col]) + 1) % P
        p2, _ = om.prove_shard({**case.traces, name: bad}, case.public_values)
        assert not om.verify_shard(p2)[0], (name, row, col)


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
@pytest.mark.parametrize("n,log_h,col_major,on_device", [(5000, 13, False, False), (5000, 13, True, True), (1025, 11, True, False),
                                                         (33, 6, False, False), (1, 4, False, True), (0, 4, True, False)])
def test_gpu_global_trace_matches_oracle(gpu, oracle, n, log_h, col_major, on_device):
    torch, prover = gpu
    ev = tg.synthetic_global_events(n, seed=20 + n)
    w, h = tg.GLOBAL_WIDTH, 1 << log_h
    out = torch.full((h * w,), -1, dtype=torch.int32, device="cuda")
    src = torch.from_numpy(ev.view(np.int32)).cuda() if on_device and n else ev
    prover.generate_alu_trace("Global", src, log_h, out, col_major=col_major)
    got = out.cpu().numpy().view(np.uint32)
    got = got.reshape(w, h).T if col_major else got.reshape(h, w)
    assert np.array_equal(got, kb.to_monty(oracle.global_trace(ev, h)))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["rows_host", "events_host", "events_device"])
def test_global_shard_proves_bit_exact(gpu, oracle, mode):
    """A shard with the Global table under the chip's restated constraints: uploaded rows, or the GlobalLookupEvent records
    handed to zkb200_commit (lift, scan and accumulation run inside the commit) - the proof is the oracle's proof over the
    oracle's rows, word for word, and the oracle's verifier accepts it."""
    from ziren_b200 import synthetic
    from ziren_b200.prover import B200Prover, EventTrace
    torch, _ = gpu
    ev = tg.synthetic_global_events(1500, seed=31)
    log_h = tg.padded_log_height(len(ev))
    case = synthetic.global_case(oracle.global_trace(ev, 1 << log_h))
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    want, _ = om.prove_shard(case.traces, case.public_values)
    prover = B200Prover(case.machine, device=0)
    try:
        pk = prover.setup({})
        inputs = {k: kb.to_monty(v) for k, v in case.traces.items()}
        if mode != "rows_host":
            d = torch.from_numpy(ev.view(np.int32)).cuda() if mode == "events_device" else ev
            inputs["Global"] = EventTrace(d, log_h, tg.GLOBAL_WIDTH)
        got, _ = prover.prove_shard(pk, inputs, case.public_values)
        ok, err = om.verify_shard(got)
        assert ok, err
        assert np.array_equal(got, want)
        pk.free()
    finally:
        prover.close()


@pytest.mark.gpu
@pytest.mark.parametrize("first_shard", [False, True])
def test_memory_and_global_tables_prove_from_event_records(gpu, oracle, first_shard):
    """The three tables that carry memory across shards handed to zkb200_commit as EVENT RECORDS (the memory events with their
    neighbour's address folded in, the global lookups they emit): rows, permutation traces of the lookup between them,
    quotients of the restated AIRs - the proof is the oracle's, word for word."""
    from ziren_b200.prover import B200Prover, EventTrace
    case, _, events = _memory_shard(oracle, first_shard, n_init=900, n_fin=2000, log_init=10, log_fin=11, seed=5,
                                    n_syscalls=0 if first_shard else 300)
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    want, _ = om.prove_shard(case.traces, case.public_values)
    prover = B200Prover(case.machine, device=0)
    try:
        pk = prover.setup({})
        inputs = {name: EventTrace(ev, log_h, tg.width(name)) for name, (ev, log_h) in events.items()}
        got, _ = prover.prove_shard(pk, inputs, case.public_values)
        ok, err = om.verify_shard(got)
        assert ok, err
        assert np.array_equal(got, want)
        pk.free()
    finally:
        prover.close()


@pytest.mark.gpu
def test_precompile_shard_tables_prove_from_event_records(gpu, oracle):
    from ziren_b200.prover import B200Prover, EventTrace
    case, _, events = _precompile_shard(oracle, 700, seed=8)
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    want, _ = om.prove_shard(case.traces, case.public_values)
    prover = B200Prover(case.machine, device=0)
    try:
        pk = prover.setup({})
        inputs = {name: EventTrace(ev, log_h, tg.width(name)) for name, (ev, log_h) in events.items()}
        got, _ = prover.prove_shard(pk, inputs, case.public_values)
        ok, err = om.verify_shard(got)
        assert ok, err
        assert np.array_equal(got, want)
        pk.free()
    finally:
        prover.close()

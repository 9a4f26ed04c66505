"""K7 (csrc/derive.cuh, zkb200_derive_multiplicities): the multiplicity columns of the receive-only tables (Byte, Program)
derived from the rows of the tables that send to them.  Reference: the host-side histograms of
crates/core/machine/src/bytes/trace.rs:46-67 (record.byte_lookups -> ByteChip::generate_trace) and
program/mod.rs:115-158 (instruction counts -> ProgramChip::generate_trace).

Three layers, as for the other row fillers:
  * the oracle's restatement (a std::map over the machine description's lookups) against the multiplicities the shard
    generators count on their own while they make the senders' rows (np.bincount in synthetic.build_case, keccak_air);
  * the product's device functions walked on the host in kernel order (tests/hostcheck) against the oracle;
  * the CUDA kernels through the C ABI against the oracle (-m gpu; sorted last: written after the round's GPU budget was
    spent, NEVER RUN ON A GPU yet).
"""
import ctypes as C

import numpy as np
import pytest

from ziren_b200 import field as kb
from ziren_b200 import synthetic

P = kb.P


def _senders(case, receiver):
    return [(n, case.prep.get(n), t) for n, t in case.traces.items() if n != receiver]


def _cases():
    yield "mini", synthetic.mini_case()
    yield "edge", synthetic.edge_case()
    yield "core", synthetic.fibonacci_core_case(log_cpu=8, num_queries=4, pow_bits=2)


@pytest.mark.parametrize("which", ["mini", "edge", "core"])
def test_oracle_derivation_reproduces_the_generators_histograms(oracle, which):
    """Byte and Program multiplicities re-derived from the senders' rows == the histograms counted while the rows were made."""
    case = dict(_cases())[which]
    om = oracle.OracleMachine(case.machine)
    for receiver in ("Byte", "Program"):
        if receiver not in case.traces:
            continue
        want = case.traces[receiver]
        got, n = om.derive_multiplicities(receiver, case.prep[receiver], _senders(case, receiver), want.shape[1])
        assert np.array_equal(got, want), receiver
        assert n > 0 and n >= int(np.count_nonzero(want))


def _keccak_case(oracle):
    from ziren_b200 import keccak_sponge as ks
    b = ks.synthetic_blocks(3, [1, 2, 1], seed=5, shard=1)
    t = oracle.keccak_sponge_trace(b, 1 << ks.padded_log_height(len(b)))
    return synthetic.keccak_real_case(b, t, log_cpu=10, num_queries=4, pow_bits=2)


def test_derivation_of_the_real_keccak_shards_byte_table(oracle, host):
    """The Byte table of ziren_b200/keccak_air.py (65536 rows, several multiplicity columns: range checks, XOR, ...) from
    the real KeccakSponge rows (3531 columns, 357 lookups) and the tables beside them: oracle == the generator's own count
    (keccak_air.receiver_tables) == the product's device functions walked on the host."""
    case = _keccak_case(oracle)
    om = oracle.OracleMachine(case.machine)
    want = case.traces["Byte"]
    got, n = om.derive_multiplicities("Byte", case.prep["Byte"], _senders(case, "Byte"), want.shape[1])
    assert np.array_equal(got, want)
    got_h, counted, misses = _host_derive(host, case, "Byte")
    assert misses == 0 and counted == n
    assert np.array_equal(got_h, want)


# ---- the product's device functions on the host ---------------------------------------------------------------------------
def _m(x):
    return (int(x) % P << 32) % P


def _flatten(machine):
    """MachineInfo::upload's DevLookup / DevVPC / DevTerm tables (csrc/machine.cpp): chips in machine order, sends then
    receives; returns (lookups n x 5, vpcs n x 3, terms n x 2, {chip: first lookup index})."""
    lookups, vpcs, terms, begin = [], [], [], {}

    def add_vpc(v):
        c, ts = v
        tb = len(terms)
        for is_main, col, w in ts:
            terms.append((col | (0x80000000 if is_main else 0), _m(w)))
        vpcs.append((_m(c), tb, len(terms)))
        return len(vpcs) - 1

    for chip in machine.chips:
        begin[chip.name] = len(lookups)
        b = chip.builder
        for is_send, l in [(1, x) for x in b.sends] + [(0, x) for x in b.receives]:
            m = add_vpc(l["mult"])
            vb = len(vpcs)
            for v in l["values"]:
                add_vpc(v)
            lookups.append((_m(l["kind"]), is_send, m, vb, len(vpcs)))
    arr = lambda x, w: np.asarray(x, dtype=np.uint32).reshape(-1, w)
    return arr(lookups, 5), arr(vpcs, 3), arr(terms, 2), begin


def _derivable(l):
    c, ts = l["mult"]
    return (l["scope"] == 0 and all(not is_main for _, vt in l["values"] for is_main, _, _ in vt) and c == 0 and len(ts) == 1 and ts[0][0] == 1
            and ts[0][2] == 1)


def _host_derive(host, case, receiver):
    machine = case.machine
    L, V, T, begin = _flatten(machine)
    rc = machine.chip(receiver)
    recv, kinds = [], set()
    for i, l in enumerate(rc.builder.receives):
        if _derivable(l):
            recv.append((begin[receiver] + len(rc.builder.sends) + i, l["mult"][1][0][1]))
            kinds.add(l["kind"])
    assert recv
    recv = np.asarray(recv, dtype=np.uint32)
    names = [n for n in case.traces if n != receiver]
    cm = lambda a: np.ascontiguousarray(kb.to_monty(np.ascontiguousarray(np.asarray(a).T)))
    preps = [cm(case.prep[n]) if n in case.prep else None for n in names]
    mains = [cm(case.traces[n]) for n in names]
    heights = [case.traces[n].shape[0] for n in names]
    send_lookup, send_table = [], []
    for ti, n in enumerate(names):
        for i, l in enumerate(machine.chip(n).builder.sends):
            if l["kind"] in kinds and l["scope"] == 0:
                send_lookup.append(begin[n] + i)
                send_table.append(ti)
    rprep = cm(case.prep[receiver])
    h = case.prep[receiver].shape[0]
    out = np.zeros((rc.main_width, h), dtype=np.uint32)
    stats = (C.c_ulonglong * 2)()
    k = len(names)
    pp = (C.c_void_p * max(1, k))(*[(p.ctypes.data if p is not None else None) for p in preps])
    mp = (C.c_void_p * max(1, k))(*[m.ctypes.data for m in mains])
    hs = (C.c_size_t * max(1, k))(*heights)
    sl = np.asarray(send_lookup, dtype=np.uint32)
    st = np.asarray(send_table, dtype=np.uint32)
    vp = lambda a: C.c_void_p(a.ctypes.data)
    host.hostcheck_derive.restype = C.c_int
    assert host.hostcheck_derive(vp(L), vp(V), vp(T), vp(recv), C.c_uint32(len(recv)), vp(rprep), C.c_size_t(h), C.c_uint32(rc.main_width),
                                 C.c_int(len(sl)), vp(sl), vp(st), pp, mp, hs, vp(out), stats) == 0
    return kb.from_monty(out).T, int(stats[0]), int(stats[1])


@pytest.mark.parametrize("which", ["mini", "edge", "core"])
def test_device_functions_on_the_host_match_the_oracle(oracle, host, which):
    case = dict(_cases())[which]
    om = oracle.OracleMachine(case.machine)
    for receiver in ("Byte", "Program"):
        if receiver not in case.traces:
            continue
        want, n = om.derive_multiplicities(receiver, case.prep[receiver], _senders(case, receiver), case.traces[receiver].shape[1])
        got, counted, misses = _host_derive(host, case, receiver)
        assert misses == 0
        assert counted == n
        assert np.array_equal(got, want), receiver
        assert np.array_equal(got, case.traces[receiver]), receiver


def _with_a_lookup_outside_the_table(case):
    """one range-checked cell of the first wide table pushed out of the Byte table's 16 bits"""
    name = next(n for n in case.traces if n not in ("Byte", "Program"))
    bad = {k: v.copy() for k, v in case.traces.items()}
    bad[name][1, 5] = 1 << synthetic.RANGE_BITS
    return synthetic.ShardCase(case.machine, case.prep, bad, case.public_values, case.cycles)


def test_a_lookup_that_is_in_no_row_is_reported(oracle, host):
    case = _with_a_lookup_outside_the_table(synthetic.mini_case())
    om = oracle.OracleMachine(case.machine)
    with pytest.raises(RuntimeError, match="in no row"):
        om.derive_multiplicities("Byte", case.prep["Byte"], _senders(case, "Byte"), 1)
    _, _, misses = _host_derive(host, case, "Byte")
    assert misses == 1


def test_repeated_tuples_go_to_the_first_row_that_holds_them(oracle, host):
    """A receiver whose preprocessed table repeats a tuple: the lowest row takes the whole multiplicity (the reference's
    fixed tables hold every tuple once, so this only fixes what the product does with a degenerate description)."""
    case = synthetic.mini_case()
    prep = dict(case.prep)
    p = prep["Byte"].copy()
    p[7] = p[3]                                                   # row 7 now repeats row 3's tuple; value 7 is in no row
    prep["Byte"] = p
    traces = {k: v.copy() for k, v in case.traces.items()}
    for n, t in traces.items():
        if n not in ("Byte", "Program"):
            chip = case.machine.chip(n)
            for l in chip.builder.sends:
                if l["kind"] == case.machine.chip("Byte").builder.receives[0]["kind"]:
                    for _, vt in l["values"]:
                        for is_main, col, _ in vt:
                            t[:, col][t[:, col] == 7] = 3
    c2 = synthetic.ShardCase(case.machine, prep, traces, case.public_values, case.cycles)
    om = oracle.OracleMachine(c2.machine)
    want, n = om.derive_multiplicities("Byte", p, _senders(c2, "Byte"), 1)
    assert want[7, 0] == 0
    got, counted, misses = _host_derive(host, c2, "Byte")
    assert misses == 0 and counted == n
    assert np.array_equal(got, want)


# ---- the CUDA kernels through the C ABI --------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("which", ["mini", "edge", "core"])
def test_derive_multiplicities_matches_oracle(oracle, which):
    import torch
    from ziren_b200.prover import B200Prover
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    case = dict(_cases())[which]
    om = oracle.OracleMachine(case.machine)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(kb.to_monty(np.ascontiguousarray(np.asarray(a).T))).view(np.int32)).cuda()
    prover = B200Prover(case.machine, device=0)
    try:
        for receiver in ("Byte", "Program"):
            if receiver not in case.traces:
                continue
            want, n = om.derive_multiplicities(receiver, case.prep[receiver], _senders(case, receiver), case.traces[receiver].shape[1])
            snd = [(name, dev(prep) if prep is not None else None, dev(tr), tr.shape[0]) for name, prep, tr in _senders(case, receiver)]
            h, w = case.traces[receiver].shape
            d_out = torch.zeros((w, h), dtype=torch.int32, device="cuda")
            got_n = prover.derive_multiplicities(receiver, dev(case.prep[receiver]), h, snd, d_out)
            prover.sync()
            assert got_n == n
            assert np.array_equal(kb.from_monty(d_out.cpu().numpy().view(np.uint32)).T, want), receiver
    finally:
        prover.close()


@pytest.mark.gpu
def test_derive_multiplicities_reports_a_missing_tuple(oracle):
    import torch
    from ziren_b200.prover import B200Prover
    case = _with_a_lookup_outside_the_table(synthetic.mini_case())
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(kb.to_monty(np.ascontiguousarray(np.asarray(a).T))).view(np.int32)).cuda()
    prover = B200Prover(case.machine, device=0)
    try:
        snd = [(name, dev(prep) if prep is not None else None, dev(tr), tr.shape[0]) for name, prep, tr in _senders(case, "Byte")]
        h, w = case.traces["Byte"].shape
        d_out = torch.zeros((w, h), dtype=torch.int32, device="cuda")
        with pytest.raises(Exception, match="in no row"):
            prover.derive_multiplicities("Byte", dev(case.prep["Byte"]), h, snd, d_out)
    finally:
        prover.close()


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["mini", "edge"])
def test_shard_proves_with_derived_tables(oracle, which):
    """ZKB200_TRACE_DERIVED: Byte and Program are not handed over at all - zkb200_prove_shard counts their multiplicity columns
    from the shard's other tables (made resident first) and the proof is the oracle's proof over the full set of tables, word
    for word; zkb200_commit, which has no proving key, refuses the flag."""
    from ziren_b200.prover import B200Prover, DerivedTrace, ZkbError
    case = dict(_cases())[which]
    om = oracle.OracleMachine(case.machine)
    om.setup(case.prep)
    want, _ = om.prove_shard(case.traces, case.public_values)
    prover = B200Prover(case.machine, device=0)
    try:
        pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
        inputs = {k: kb.to_monty(v) for k, v in case.traces.items()}
        for r in ("Byte", "Program"):
            if r in inputs:
                inputs[r] = DerivedTrace(*case.traces[r].shape)
        got, _ = prover.prove_shard(pk, inputs, case.public_values)
        ok, err = om.verify_shard(got)
        assert ok, err
        assert np.array_equal(got, want)
        with pytest.raises(ZkbError, match="bad zkb200_trace.flags"):
            prover.commit(inputs, case.public_values)
        pk.free()
    finally:
        prover.close()


@pytest.mark.gpu
def test_real_keccak_shard_proves_from_events_with_a_derived_byte_table(oracle):
    """The shard of the real KeccakSponge chip (357 lookups per row) with the chip's table handed over as EVENT RECORDS and the
    65536-row Byte table not handed over at all: rows from the row filler, Byte multiplicities counted from them and from the
    other tables inside zkb200_prove_shard - the oracle's proof over the oracle's tables, word for word."""
    from ziren_b200 import keccak_sponge as ks
    from ziren_b200.prover import B200Prover, DerivedTrace, EventTrace
    b = ks.synthetic_blocks(3, [1, 2, 1], seed=5, shard=1)
    log_h = ks.padded_log_height(len(b))
    case = synthetic.keccak_real_case(b, oracle.keccak_sponge_trace(b, 1 << log_h), log_cpu=10, num_queries=4, pow_bits=2)
    om = oracle.OracleMachine(case.machine)
    om.setup(case.prep)
    want, _ = om.prove_shard(case.traces, case.public_values)
    prover = B200Prover(case.machine, device=0)
    try:
        pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
        inputs = {k: kb.to_monty(v) for k, v in case.traces.items()}
        inputs["KeccakSponge"] = EventTrace(b.reshape(len(b), -1), log_h, ks.WIDTH)
        inputs["Byte"] = DerivedTrace(*case.traces["Byte"].shape)
        got, _ = prover.prove_shard(pk, inputs, case.public_values)
        ok, err = om.verify_shard(got)
        assert ok, err
        assert np.array_equal(got, want)
        pk.free()
    finally:
        prover.close()

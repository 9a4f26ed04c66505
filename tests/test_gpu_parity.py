"""GPU parity tests: the CUDA path, called through the C ABI (libzkb200.so), against the CPU oracle
on the same seeded inputs.  Integer work: every comparison is bit-exact."""
import numpy as np
import pytest

from ziren_b200 import field as kb
from ziren_b200 import synthetic

pytestmark = pytest.mark.gpu
P = kb.P


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def mini(torch):
    from ziren_b200.prover import B200Prover
    case = synthetic.mini_case()
    prover = B200Prover(case.machine, device=0)
    yield case, prover
    prover.close()


def dev(torch, a):
    """uint32 numpy -> CUDA int32 tensor with the same bits"""
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint32)


def colmajor(a):
    return np.ascontiguousarray(np.asarray(a).T)


def test_poseidon2_permute_batch(torch, mini, oracle):
    _, prover = mini
    rng = np.random.default_rng(1)
    st = kb.random_elements(rng, (1000, 16))
    d = dev(torch, kb.to_monty(st))
    prover.poseidon2_permute_batch(d, 1000)
    prover.sync()
    assert np.array_equal(kb.from_monty(host(d)), oracle.permute_batch(st))
    # the reference's known-answer vector: perm([0..15])
    d = dev(torch, kb.to_monty(np.arange(16, dtype=np.uint32)).reshape(1, 16))
    prover.poseidon2_permute_batch(d, 1)
    prover.sync()
    assert kb.from_monty(host(d))[0, :4].tolist() == [1635930443, 1105042214, 1882043429, 1844048402]


@pytest.mark.parametrize("log_n,width", [(1, 3), (2, 1), (4, 5), (7, 3), (10, 9), (13, 2), (16, 2), (17, 1)])
def test_ntt_matches_oracle(torch, mini, oracle, log_n, width):
    _, prover = mini
    rng = np.random.default_rng(log_n * 100 + width)
    x = kb.random_elements(rng, (1 << log_n, width))
    for inverse in (False, True):
        want = oracle.dft(x, inverse=inverse)
        d_in = dev(torch, kb.to_monty(colmajor(x)))
        d_out = torch.empty_like(d_in)
        prover.ntt(d_in, d_out, log_n, width, inverse=inverse, bitrev_out=False)
        prover.sync()
        got = kb.from_monty(host(d_out)).T
        assert np.array_equal(got, want), f"log_n={log_n} inverse={inverse}"


@pytest.mark.parametrize("log_n,width,log_blowup,shift", [(1, 2, 1, 3), (3, 4, 1, 3), (5, 1, 2, 3), (8, 7, 1, 3), (9, 3, 1, 5),
                                                          (12, 4, 1, 3), (15, 2, 1, 3), (16, 3, 2, 3), (18, 1, 1, 3)])
def test_coset_lde_matches_oracle(torch, mini, oracle, log_n, width, log_blowup, shift):
    _, prover = mini
    rng = np.random.default_rng(log_n * 1000 + width)
    x = kb.random_elements(rng, (1 << log_n, width))
    want = oracle.coset_lde(x, added_bits=log_blowup, shift=shift)
    d_in = dev(torch, kb.to_monty(colmajor(x)))
    d_out = torch.empty((width, (1 << log_n) << log_blowup), dtype=torch.int32, device="cuda")
    prover.coset_lde(d_in, d_out, log_n, width, log_blowup, shift)
    prover.sync()
    assert np.array_equal(kb.from_monty(host(d_out)).T, want)


def test_lde_roundtrip_property_large(torch, mini):
    """size-independent property at a size the oracle would not finish quickly: the even rows of
    the natural-order LDE on shift 1 are the original evaluations (2^20 x 4)."""
    _, prover = mini
    log_n, width = 20, 4
    rng = np.random.default_rng(7)
    x = kb.random_elements(rng, (width, 1 << log_n))
    d_in = dev(torch, kb.to_monty(x))
    d_out = torch.empty((width, 2 << log_n), dtype=torch.int32, device="cuda")
    prover.coset_lde(d_in, d_out, log_n, width, 1, 1)     # shift 1: the LDE domain contains H_n
    prover.sync()
    got = kb.from_monty(host(d_out))
    # bit-reversed storage: first n rows = coset w_{2n}^even = H_n in bit-reversed order
    n = 1 << log_n
    idx = np.arange(n, dtype=np.uint32)
    rev = np.zeros(n, dtype=np.uint32)
    for b in range(log_n):
        rev |= ((idx >> b) & 1) << (log_n - 1 - b)
    assert np.array_equal(got[:, :n], x[:, rev])


@pytest.mark.parametrize("shapes", [[(4, 3)], [(6, 9), (6, 1), (4, 17), (3, 2)], [(10, 8), (10, 16), (9, 5), (2, 1)],
                                    [(12, 33), (5, 0), (5, 3)]])
def test_mmcs_root_matches_oracle(torch, mini, oracle, shapes):
    _, prover = mini
    rng = np.random.default_rng(len(shapes))
    mats = [kb.random_elements(rng, (1 << lh, w)) for lh, w in shapes]
    want = oracle.mmcs_root(mats)
    dmats = [dev(torch, kb.to_monty(colmajor(m))) if m.size else torch.empty(1, dtype=torch.int32, device="cuda") for m in mats]
    got = prover.mmcs_root(dmats, [lh for lh, _ in shapes], [w for _, w in shapes])
    assert np.array_equal(got, want)


def test_commit_matches_oracle(torch, mini, oracle):
    case, prover = mini
    om = oracle.OracleMachine(case.machine)
    want_prep = om.setup(case.prep)
    pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
    assert np.array_equal(pk.commit, want_prep)
    assert np.array_equal(pk.observe_into(), om.initial_challenger())
    data = prover.commit({k: kb.to_monty(v) for k, v in case.traces.items()}, case.public_values)
    assert np.array_equal(data.main_commit, om.commit_shard(case.traces))
    data.free()
    pk.free()


@pytest.mark.parametrize("codegen", [1, 0], ids=["generated", "data-driven"])
def test_permutation_trace_matches_oracle(torch, mini, oracle, codegen):
    """K5 both ways: the rows from the chip's generated module (kernel `lk`) and from the data-driven kernel over the
    flattened lookups, each bit-exact against the oracle for every chip of the machine."""
    from ziren_b200 import _ffi
    _ffi.lib().zkb200_set_option(b"logup_codegen", codegen)
    try:
        _permutation_trace_matches_oracle(torch, mini, oracle)
    finally:
        _ffi.lib().zkb200_set_option(b"logup_codegen", 1)


def _permutation_trace_matches_oracle(torch, mini, oracle):
    case, prover = mini
    om = oracle.OracleMachine(case.machine)
    rng = np.random.default_rng(3)
    alpha, beta = kb.random_elements(rng, 4), kb.random_elements(rng, 4)
    for name, tr in case.traces.items():
        prep = case.prep.get(name)
        want, want_sum = om.permutation_trace(name, prep, tr, alpha, beta)
        ew = case.machine.chip(name).perm_width_ef
        d_main = dev(torch, kb.to_monty(colmajor(tr)))
        d_prep = dev(torch, kb.to_monty(colmajor(prep))) if prep is not None else None
        d_out = torch.zeros((4 * ew, tr.shape[0]), dtype=torch.int32, device="cuda")
        got_sum = prover.permutation_trace(name, d_prep, d_main, tr.shape[0], alpha, beta, d_out)
        prover.sync()
        assert np.array_equal(kb.from_monty(host(d_out)).T, want), name
        assert np.array_equal(got_sum, want_sum), name


@pytest.mark.parametrize("codegen", [1, 0], ids=["generated", "interpreter"])
def test_quotient_matches_oracle(torch, mini, oracle, codegen):
    """K3 both ways: the per-chip kernels generated and compiled at run time (NVRTC) and the bytecode
    interpreter they replace, each bit-exact against the oracle for every chip of the machine."""
    from ziren_b200 import _ffi
    import ctypes as C
    _ffi.lib().zkb200_set_option(b"quotient_codegen", codegen)
    g0, i0 = C.c_ulonglong(), C.c_ulonglong()
    _ffi.lib().zkb200_quotient_launch_counts(C.byref(g0), C.byref(i0))
    try:
        _quotient_matches_oracle(torch, mini, oracle)
    finally:
        _ffi.lib().zkb200_set_option(b"quotient_codegen", 1)
    g1, i1 = C.c_ulonglong(), C.c_ulonglong()
    _ffi.lib().zkb200_quotient_launch_counts(C.byref(g1), C.byref(i1))
    # the path under test is the one that ran (a generated kernel that failed to build would fall back silently)
    if codegen:
        assert g1.value > g0.value and i1.value == i0.value
    else:
        assert i1.value > i0.value and g1.value == g0.value


def _quotient_matches_oracle(torch, mini, oracle):
    case, prover = mini
    om = oracle.OracleMachine(case.machine)
    rng = np.random.default_rng(4)
    pa, pb, al = (kb.random_elements(rng, 4) for _ in range(3))
    for name, tr in case.traces.items():
        chip = case.machine.chip(name)
        prep = case.prep.get(name)
        log_n = int(np.log2(tr.shape[0]))
        perm, lsum = om.permutation_trace(name, prep, tr, pa, pb)
        gsum = tr[-1, -14:] if chip.global_scope else kb.SEPTIC_DIGEST_ZERO
        main_lde = oracle.coset_lde(tr)
        perm_lde = oracle.coset_lde(perm) if perm.shape[1] else np.zeros((2 * tr.shape[0], 0), np.uint32)
        prep_lde = oracle.coset_lde(prep) if prep is not None else None
        want = om.quotient_values(name, log_n, prep_lde, main_lde, perm_lde, pa, pb, lsum, gsum, al, case.public_values)
        nch = 1 << chip.log_quotient_degree
        n = tr.shape[0]
        d_out = torch.zeros((nch, 4, n), dtype=torch.int32, device="cuda")
        d = lambda a: dev(torch, kb.to_monty(colmajor(a))) if a is not None and a.size else (None if a is None else torch.empty(1, dtype=torch.int32, device="cuda"))
        prover.quotient(name, log_n, d(prep_lde), d(main_lde), d(perm_lde), pa, pb, lsum, gsum, al, case.public_values, d_out)
        got = kb.from_monty(host(d_out))           # [chunk][c][k]
        # oracle: natural order Q x 4, q[k*nch + j] = chunk j row k
        want_chunks = want.reshape(n, nch, 4).transpose(1, 2, 0)
        assert np.array_equal(got, want_chunks), name


def test_fri_fold_matches_oracle(torch, mini, oracle):
    _, prover = mini
    rng = np.random.default_rng(5)
    for log_m in (2, 5, 11):
        m = 1 << log_m
        vals = kb.random_elements(rng, (m, 4))
        beta = kb.random_elements(rng, 4)
        ro = kb.random_elements(rng, (m // 2, 4))
        for r in (None, ro):
            want = oracle.fri_fold(vals, beta, r)
            d_in = dev(torch, kb.to_monty(colmajor(vals)))
            d_ro = dev(torch, kb.to_monty(colmajor(r))) if r is not None else None
            d_out = torch.empty((4, m // 2), dtype=torch.int32, device="cuda")
            prover.fri_fold(d_in, m, beta, d_ro, d_out)
            prover.sync()
            assert np.array_equal(kb.from_monty(host(d_out)).T, want)


def test_grind_matches_oracle(torch, mini, oracle):
    _, prover = mini
    rng = np.random.default_rng(6)
    st = np.zeros(34, np.uint32)
    st[:16] = kb.random_elements(rng, 16)
    for n_in in (0, 3, 7):
        s = st.copy()
        s[16] = n_in
        s[17:17 + n_in] = kb.random_elements(rng, n_in)
        _, want = oracle.grind(s, 12)
        assert prover.grind(s, 12) == want


def test_shard_proof_bit_exact_and_verifies(torch, mini, oracle):
    case, prover = mini
    om = oracle.OracleMachine(case.machine)
    om.setup(case.prep)
    want, want_ch = om.prove_shard(case.traces, case.public_values)
    pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
    got, got_ch = prover.prove_shard(pk, {k: kb.to_monty(v) for k, v in case.traces.items()}, case.public_values)
    ok, err = om.verify_shard(got)
    assert ok, err
    assert got.size == want.size
    assert np.array_equal(got, want)
    assert np.array_equal(got_ch, want_ch)
    pk.free()


def test_device_resident_inputs_give_same_proof(torch, mini, oracle):
    case, prover = mini
    pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
    host_tr = {k: kb.to_monty(v) for k, v in case.traces.items()}
    dev_tr = {k: dev(torch, v) for k, v in host_tr.items()}
    a, _ = prover.prove_shard(pk, host_tr, case.public_values)
    b, _ = prover.prove_shard(pk, dev_tr, case.public_values)
    assert np.array_equal(a, b)
    pk.free()


@pytest.mark.parametrize("seed,log_cpu", [(2, 8), (3, 11)])
def test_fibonacci_core_shard_bit_exact(torch, oracle, seed, log_cpu):
    from ziren_b200.prover import B200Prover
    case = synthetic.fibonacci_core_case(log_cpu=log_cpu, seed=seed, num_queries=6, pow_bits=8)
    prover = B200Prover(case.machine)
    om = oracle.OracleMachine(case.machine)
    om.setup(case.prep)
    want, _ = om.prove_shard(case.traces, case.public_values)
    pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
    got, _ = prover.prove_shard(pk, {k: kb.to_monty(v) for k, v in case.traces.items()}, case.public_values)
    assert np.array_equal(got, want)
    ok, err = om.verify_shard(got)
    assert ok, err
    pk.free()
    prover.close()


@pytest.mark.parametrize("which", ["edge", "edge_blowup2", "mini_blowup2", "mini_blowup3", "noprep"])
def test_edge_machines_bit_exact(torch, oracle, which):
    """chips without lookups (log_quotient_degree 0, zero-width permutation matrices), 2-row tables,
    a machine without any preprocessed table (empty proving-key commitment),
    FRI blow-up 4 and 8 (the reference's compressed()/ultra_compressed() configs,
    crates/stark/src/kb31_poseidon2.rs:217-241)."""
    from ziren_b200.prover import B200Prover
    case = {"edge": lambda: synthetic.edge_case(),
            "edge_blowup2": lambda: synthetic.edge_case(log_blowup=2),
            "mini_blowup2": lambda: synthetic.mini_case(seed=9, log_blowup=2, num_queries=5),
            "mini_blowup3": lambda: synthetic.mini_case(seed=10, log_blowup=3, num_queries=4),
            "noprep": lambda: synthetic.noprep_case()}[which]()
    prover = B200Prover(case.machine)
    om = oracle.OracleMachine(case.machine)
    om.setup(case.prep)
    want, _ = om.prove_shard(case.traces, case.public_values)
    pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
    got, _ = prover.prove_shard(pk, {k: kb.to_monty(v) for k, v in case.traces.items()}, case.public_values)
    ok, err = om.verify_shard(got)
    assert ok, err
    assert np.array_equal(got, want)
    pk.free()
    prover.close()


def test_concurrent_shards_are_independent(torch, mini, oracle):
    """several host threads proving different shards on one context (compute lanes + copy stream)
    give the same proofs as sequential proving"""
    import threading
    case, prover = mini
    pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
    cases = [synthetic.mini_case(seed=20 + i) for i in range(4)]
    traces = [{k: kb.to_monty(v) for k, v in c.traces.items()} for c in cases]
    seq = [prover.prove_shard(pk, tr, c.public_values)[0] for tr, c in zip(traces, cases)]
    out = [None] * len(cases)

    def work(i):
        out[i] = prover.prove_shard(pk, traces[i], cases[i].public_values)[0]
    ths = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for a, b in zip(seq, out):
        assert np.array_equal(a, b)
    pk.free()


@pytest.mark.parametrize("which", ["core", "compress", "keccak"])
def test_reference_shapes_bit_exact(torch, oracle, which):
    """the other BASELINE.json configs as parity cases at a size the oracle proves in seconds:
    tendermint-like maximal core shape (18 execution tables + Byte + Program), recursion-compress
    shape (Poseidon2Wide 313 columns), keccak-precompile shape (4167-column table)"""
    from ziren_b200.prover import B200Prover
    case = {"core": lambda: synthetic.core_case(log_cpu=9, seed=31, num_queries=5, pow_bits=6),
            "compress": lambda: synthetic.compress_case(log_max=9, seed=32, num_queries=5, pow_bits=6),
            "keccak": lambda: synthetic.keccak_case(log_cpu=8, seed=33, num_queries=5, pow_bits=6)}[which]()
    prover = B200Prover(case.machine)
    om = oracle.OracleMachine(case.machine)
    om.setup(case.prep)
    want, _ = om.prove_shard(case.traces, case.public_values)
    pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
    got, _ = prover.prove_shard(pk, {k: kb.to_monty(v) for k, v in case.traces.items()}, case.public_values)
    assert np.array_equal(got, want)
    ok, err = om.verify_shard(got)
    assert ok, err
    pk.free()
    prover.close()


def test_large_shard_is_accepted_by_the_verifier(torch, oracle):
    """Size-independent property at a size the CPU oracle cannot re-prove in test time: a
    keccak-precompile-like shard (Cpu 2^17 rows, KeccakSponge 2^15 x 4167 columns, 168 M cells, the
    reference's FRI parameters: 84 queries, 16 PoW bits) proved on the GPU must be ACCEPTED by the
    oracle's verifier (restated from crates/stark/src/verifier.rs and the recursion circuit), and a
    corrupted copy must be rejected."""
    from ziren_b200.prover import B200Prover
    case = synthetic.keccak_case(log_cpu=17, seed=123)
    prover = B200Prover(case.machine)
    om = oracle.OracleMachine(case.machine)
    commit_want = None
    pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
    proof, _ = prover.prove_shard(pk, {k: kb.to_monty(v) for k, v in case.traces.items()}, case.public_values)
    om.setup(case.prep)       # CPU: preprocessed commit only (2^17 x 3 cells)
    ok, err = om.verify_shard(proof)
    assert ok, err
    bad = proof.copy()
    bad[proof.size // 2] ^= 1
    assert not om.verify_shard(bad)[0]
    pk.free()
    prover.close()


def test_launch_counter_counts(torch, mini):
    case, prover = mini
    n0 = prover.launch_count()
    d = dev(torch, kb.to_monty(np.arange(16, dtype=np.uint32)).reshape(1, 16))
    prover.poseidon2_permute_batch(d, 1)
    assert prover.launch_count() == n0 + 1


def test_errors_are_reported(torch, mini):
    from ziren_b200.prover import ZkbError
    case, prover = mini
    with pytest.raises(ZkbError, match="unknown chip"):
        prover.commit({"Nope": np.zeros((4, 1), np.uint32)}, case.public_values)
    with pytest.raises(ZkbError, match="width mismatch"):
        prover.commit({"Cpu": np.zeros((4, 1), np.uint32)}, case.public_values)


# ---- sizes the reference really runs (crates/stark/src/opts.rs:42-50: shards of 2^21 / 2^22 rows) -------

@pytest.mark.parametrize("log_n", [20, 22, 24])
def test_ntt_bit_exact_at_baseline_sizes(torch, mini, oracle, log_n):
    """BASELINE configs[3] sizes, bit-exact against the oracle (one column; 2^23 and 2^24 take the
    K1 = K2 = 12 split of the two-level transform)."""
    _, prover = mini
    rng = np.random.default_rng(log_n)
    x = kb.random_elements(rng, (1 << log_n, 1))
    for inverse in (False, True):
        want = oracle.dft(x, inverse=inverse)
        d_in = dev(torch, kb.to_monty(colmajor(x)))
        d_out = torch.empty_like(d_in)
        prover.ntt(d_in, d_out, log_n, 1, inverse=inverse, bitrev_out=False)
        prover.sync()
        assert np.array_equal(kb.from_monty(host(d_out)).T, want), f"log_n={log_n} inverse={inverse}"


@pytest.mark.parametrize("log_n,width,log_blowup", [(20, 2, 1), (21, 1, 1), (22, 1, 1), (23, 1, 1), (22, 1, 2)])
def test_coset_lde_bit_exact_at_baseline_sizes(torch, mini, oracle, log_n, width, log_blowup):
    _, prover = mini
    rng = np.random.default_rng(log_n * 7 + width)
    x = kb.random_elements(rng, (1 << log_n, width))
    want = oracle.coset_lde(x, added_bits=log_blowup, shift=3)
    d_in = dev(torch, kb.to_monty(colmajor(x)))
    d_out = torch.empty((width, (1 << log_n) << log_blowup), dtype=torch.int32, device="cuda")
    prover.coset_lde(d_in, d_out, log_n, width, log_blowup, 3)
    prover.sync()
    assert np.array_equal(kb.from_monty(host(d_out)).T, want)


def test_coset_lde_multi_chunk_bit_exact(torch, mini, oracle):
    """width * n > 2^26: the LDE walks the columns in chunks (ntt.cu: coset_lde_batch); the columns
    either side of every chunk boundary are compared with the oracle (columns are independent)."""
    _, prover = mini
    log_n, width = 16, 1100          # chunk = 2^26 >> 16 = 1024 columns
    rng = np.random.default_rng(99)
    x = kb.random_elements(rng, (width, 1 << log_n))          # column-major
    d_in = dev(torch, kb.to_monty(x))
    d_out = torch.empty((width, 2 << log_n), dtype=torch.int32, device="cuda")
    prover.coset_lde(d_in, d_out, log_n, width, 1, 3)
    prover.sync()
    cols = [0, 1, 511, 1022, 1023, 1024, 1025, 1099]
    got = kb.from_monty(host(d_out[cols]))
    want = oracle.coset_lde(np.ascontiguousarray(x[cols].T), added_bits=1, shift=3)
    assert np.array_equal(got.T, want)


@pytest.mark.parametrize("log_cpu", [21, 22])
def test_default_shard_sizes_prove_and_verify(torch, oracle, log_cpu):
    """The reference's default shard sizes (2^21 rows on 49-80 GB hosts, 2^22 above,
    crates/stark/src/opts.rs:42-50; BASELINE configs[2] is the maximal core shape with a 2^21-row Cpu
    table): proved on the GPU, ACCEPTED by the oracle's verifier, a corrupted copy rejected."""
    from ziren_b200.prover import B200Prover
    case = synthetic.core_case(log_cpu=log_cpu, seed=40 + log_cpu)
    prover = B200Prover(case.machine)
    om = oracle.OracleMachine(case.machine)
    pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
    traces = {}
    for k in list(case.traces):
        traces[k] = kb.to_monty(case.traces[k])
        case.traces[k] = None
    proof, _ = prover.prove_shard(pk, traces, case.public_values)
    want_prep = om.setup(case.prep)       # CPU: preprocessed commit only
    assert np.array_equal(pk.commit, want_prep)
    ok, err = om.verify_shard(proof)
    assert ok, err
    bad = proof.copy()
    bad[proof.size // 3] ^= 1
    assert not om.verify_shard(bad)[0]
    pk.free()
    prover.close()


def test_piecewise_commit_matches_oracle(torch, oracle, monkeypatch):
    """The main commit cut into many column pieces (ZKB200_PIECE_KB=16: the 4167-column table comes in
    ~260 pieces and is hashed piecewise under the upload) from pageable, pinned and device-resident
    sources: same proof, bit-exact against the oracle."""
    from ziren_b200.prover import B200Prover
    monkeypatch.setenv("ZKB200_PIECE_KB", "16")
    monkeypatch.setenv("ZKB200_STAGE_SLOT_MB", "1")
    case = synthetic.keccak_case(log_cpu=10, seed=77, num_queries=5, pow_bits=6)
    prover = B200Prover(case.machine)
    om = oracle.OracleMachine(case.machine)
    om.setup(case.prep)
    want, _ = om.prove_shard(case.traces, case.public_values)
    pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
    pageable = {k: kb.to_monty(v) for k, v in case.traces.items()}
    pinned = {k: torch.from_numpy(v.view(np.int32)).pin_memory() for k, v in pageable.items()}
    resident = {k: v.cuda() for k, v in pinned.items()}
    for name, tr in (("pageable", pageable), ("pinned", pinned), ("device", resident)):
        got, _ = prover.prove_shard(pk, tr, case.public_values)
        assert np.array_equal(got, want), name
    pk.free()
    prover.close()


def test_multi_device_prover_routes_shards(torch, oracle):
    """One prover object over every visible GPU (zkb200_ctx_create_multi): shards committed from
    several host threads are spread over the devices, every proof equals the single-GPU proof.
    On a one-GPU box the same object degenerates to one device."""
    import threading
    from ziren_b200.prover import B200Prover
    cases = [synthetic.mini_case(seed=60 + i) for i in range(6)]
    single = B200Prover(cases[0].machine, device=0)
    pk1 = single.setup({k: kb.to_monty(v) for k, v in cases[0].prep.items()})
    want = [single.prove_shard(pk1, {k: kb.to_monty(v) for k, v in c.traces.items()}, c.public_values)[0] for c in cases]
    pk1.free()
    single.close()
    multi = B200Prover(cases[0].machine, device=-1)
    assert multi.num_devices() == torch.cuda.device_count()
    pk = multi.setup({k: kb.to_monty(v) for k, v in cases[0].prep.items()})
    got, devices = [None] * len(cases), [None] * len(cases)
    gate = threading.Barrier(len(cases))

    def work(i):
        c = cases[i]
        data = multi.commit({k: kb.to_monty(v) for k, v in c.traces.items()}, c.public_values)
        devices[i] = data.device()
        gate.wait()                      # all shards are in flight before the first one is released
        got[i], _ = multi.open(pk, data)
        data.free()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    [t.start() for t in ths]
    [t.join() for t in ths]
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    if multi.num_devices() > 1:
        assert len(set(devices)) == min(multi.num_devices(), len(cases)), devices
    pk.free()
    multi.close()


@pytest.mark.parametrize("which", ["core", "keccak"])
def test_interpreter_and_generated_kernels_give_the_same_proof(torch, oracle, which):
    """whole proofs with the constraint interpreter (quotient_codegen = 0) equal the proofs made with the
    generated kernels, which the other tests compare with the oracle"""
    from ziren_b200 import _ffi
    from ziren_b200.prover import B200Prover
    case = {"core": lambda: synthetic.core_case(log_cpu=9, seed=31, num_queries=5, pow_bits=6),
            "keccak": lambda: synthetic.keccak_case(log_cpu=8, seed=33, num_queries=5, pow_bits=6)}[which]()
    prover = B200Prover(case.machine)
    pk = prover.setup({k: kb.to_monty(v) for k, v in case.prep.items()})
    tr = {k: kb.to_monty(v) for k, v in case.traces.items()}
    a, _ = prover.prove_shard(pk, tr, case.public_values)
    _ffi.lib().zkb200_set_option(b"quotient_codegen", 0)
    try:
        b, _ = prover.prove_shard(pk, tr, case.public_values)
    finally:
        _ffi.lib().zkb200_set_option(b"quotient_codegen", 1)
    assert np.array_equal(a, b)
    pk.free()
    prover.close()

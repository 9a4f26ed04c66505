"""The K3 constraint kernels the generator writes (ziren_b200/csrc/quotient_codegen.cpp: constraint shapes, parameter tables,
LogUp batch shapes, on top of the run-time header quotient_rt.cuh), compiled FOR THE HOST behind a few shims
(tests/hostcheck/qk_host.cpp) and run row by row against the oracle's quotient (oracle/air.h quotient_values) - the same
comparison tests/test_gpu_parity.py::test_quotient_matches_oracle makes on a GPU, here for the generated SOURCE without one.
Chips: the real KeccakSponge chip (3 788 constraints, 357 lookups, 19 + 8 shapes) and the chips of a small machine with
preprocessed columns, public values, the global-scope rows and lookup batches of one and two."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from ziren_b200 import _ffi
from ziren_b200 import field as kb
from ziren_b200.air import SCOPE_LOCAL

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
P = kb.P
W4 = 3          # EF = F[x] / (x^4 - 3)


def ef_mul(a, b):
    r = [0, 0, 0, 0, 0, 0, 0]
    for i in range(4):
        for j in range(4):
            r[i + j] += a[i] * b[j]
    return [(r[k] + W4 * (r[k + 4] if k + 4 < 7 else 0)) % P for k in range(4)]


def ef_add(a, b):
    return [(x + y) % P for x, y in zip(a, b)]


def ef_scale(a, s):
    return [(x * s) % P for x in a]


def two_adic_generator(bits):
    g = pow(3, 127, P)
    for _ in range(bits, 24):
        g = g * g % P
    return g


def monty(x):
    return kb.to_monty(np.asarray(x, dtype=np.uint32).reshape(-1)).reshape(np.asarray(x).shape)


@pytest.fixture(scope="module")
def tw_tables():
    w = two_adic_generator(24)
    lo = np.empty(4096, np.uint32)
    hi = np.empty(4096, np.uint32)
    x, w4096 = 1, pow(w, 4096, P)
    y = 1
    for e in range(4096):
        lo[e], hi[e] = x, y
        x, y = x * w % P, y * w4096 % P
    return monty(lo), monty(hi)


def _build_host_kernel(machine, chip_name, tmp):
    desc = np.ascontiguousarray(machine.descriptor(), dtype=np.uint32)
    n = C.c_size_t()
    os.environ["ZKB200_CODEGEN_DUMP"] = str(tmp)
    try:
        assert _ffi.lib().zkb200_codegen_compile_check(desc.ctypes.data_as(_ffi.u32p), desc.size, C.byref(n)) == len(machine.chips)
    finally:
        del os.environ["ZKB200_CODEGEN_DUMP"]
    src = os.path.join(str(tmp), f"qk_{chip_name}.cu")
    so = os.path.join(str(tmp), f"qk_{chip_name}.so")
    has_lk = ["-DQK_HAS_LK"] if " lk(const PermArgs a)" in open(src).read() else []
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-w", "-I" + os.path.join(ROOT, "ziren_b200", "csrc"),
                           f'-DQK_SOURCE="{src}"', *has_lk, os.path.join(HERE, "hostcheck", "qk_host.cpp"), "-o", so])
    return C.CDLL(so)


def _check_chip(oracle, om, machine, name, prep, trace, public_values, tw_tables, tmp, seed=4):
    chip = machine.chip(name)
    lib = _build_host_kernel(machine, name, tmp)
    rng = np.random.default_rng(seed)
    pa, pb, al = ([int(v) for v in kb.random_elements(rng, 4)] for _ in range(3))
    n = trace.shape[0]
    log_n, lqd = int(np.log2(n)), chip.log_quotient_degree
    perm, lsum = om.permutation_trace(name, prep, trace, pa, pb)
    gsum = trace[-1, -14:] if chip.global_scope else kb.SEPTIC_DIGEST_ZERO
    lb = machine.log_blowup
    main_lde = oracle.coset_lde(trace, added_bits=lb)
    perm_lde = oracle.coset_lde(perm, added_bits=lb) if perm.shape[1] else np.zeros((n << lb, 0), np.uint32)
    prep_lde = oracle.coset_lde(prep, added_bits=lb) if prep is not None else None
    want = om.quotient_values(name, log_n, prep_lde, main_lde, perm_lde, pa, pb, lsum, gsum, al, public_values)
    # what quotient.cu / lookup_coefficients prepare on the device, here in Python
    lookups = [l for l in chip.builder.sends + chip.builder.receives if l["scope"] == SCOPE_LOCAL]
    bpow = [[1, 0, 0, 0]]
    for _ in range(16):
        bpow.append(ef_mul(bpow[-1], pb))
    K, E = [], []
    for l in lookups:
        k = ef_add(pa, [l["kind"], 0, 0, 0])
        for j, (const, terms) in enumerate(l["values"], start=1):
            k = ef_add(k, ef_scale(bpow[j], const))
            for _is_main, _col, w in terms:
                E.append(ef_scale(bpow[j], w))
        K.append(k)
    ncons = chip.num_constraints
    apow = [[1, 0, 0, 0]]
    for _ in range(ncons - 1):
        apow.append(ef_mul(apow[-1], al))
    apow = apow[::-1]                                    # alpha_pow[k] = alpha^(C-1-k)
    gn = pow(3, n, P)
    wq = two_adic_generator(lqd)
    zh = [(gn * pow(wq, v, P) - 1) % P for v in range(1 << lqd)]
    inv_zh = [pow(z, P - 2, P) for z in zh]
    arr = lambda x, shape=None: np.ascontiguousarray(monty(np.array(x, dtype=np.uint32).reshape(shape or -1)))
    colmajor = lambda a: np.ascontiguousarray(a.T)
    m_prep = monty(colmajor(prep_lde)) if prep_lde is not None else np.zeros(1, np.uint32)
    m_main, m_perm = monty(colmajor(main_lde)), (monty(colmajor(perm_lde)) if perm_lde.size else np.zeros(1, np.uint32))
    a_apow, a_K, a_E = arr(apow, (-1, 4)), arr(K or [[0] * 4], (-1, 4)), arr(E or [[0] * 4], (-1, 4))
    a_pub = arr(np.asarray(public_values, dtype=np.uint64) % P)
    a_lsum, a_gsum, a_zh, a_izh = arr(lsum), arr(gsum), arr(zh), arr(inv_zh)
    tw_lo, tw_hi = tw_tables
    nch = 1 << lqd
    out = np.zeros((nch, 4, n), np.uint32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    rc = lib.qk_host_run(C.c_uint(log_n), C.c_uint(lqd), C.c_size_t(n << lb), p(m_prep), p(m_main), p(m_perm),
                         C.c_uint(chip.perm_width_ef), C.c_uint(1 << lqd), C.c_uint(chip.main_width), C.c_uint(int(chip.global_scope)),
                         C.c_uint(len(chip.builder.constraints)), C.c_uint(len(lookups)), p(a_apow), p(a_K), p(a_E), p(a_pub),
                         p(tw_lo), p(tw_hi), p(a_lsum), p(a_gsum), p(a_zh), p(a_izh),
                         C.c_uint32(int(monty(np.array([3], np.uint32))[0])),
                         C.c_uint32(int(monty(np.array([pow(two_adic_generator(log_n), P - 2, P)], np.uint32))[0])), p(out))
    assert rc == 0
    got = kb.from_monty(out.reshape(-1)).reshape(out.shape)
    want_chunks = want.reshape(n, nch, 4).transpose(1, 2, 0)
    assert np.array_equal(got, want_chunks), name
    # K5 from the same module: the permutation trace's batch columns and the row sums (the running-sum column is the
    # scan kernels' job), against the oracle's permutation trace
    if hasattr(lib, "lk_host_run") and chip.perm_width_ef:
        ew = chip.perm_width_ef
        t_prep = monty(colmajor(prep)) if prep is not None else np.zeros(1, np.uint32)
        t_main = monty(colmajor(trace))
        pout = np.zeros((4 * ew, n), np.uint32)
        rowsum = np.zeros((4, n), np.uint32)
        assert lib.lk_host_run(C.c_size_t(n), p(t_prep), p(t_main), p(a_K), p(a_E), p(pout), p(rowsum)) == 0
        got_perm = kb.from_monty(pout.reshape(-1)).reshape(pout.shape).T            # n x 4E
        assert np.array_equal(got_perm[:, :4 * (ew - 1)], perm[:, :4 * (ew - 1)]), name
        batches = perm[:, :4 * (ew - 1)].reshape(n, ew - 1, 4).astype(np.uint64).sum(axis=1) % P
        assert np.array_equal(kb.from_monty(rowsum.reshape(-1)).reshape(4, n).T.astype(np.uint64), batches), name


def test_generated_keccak_kernel_matches_the_oracle_on_the_host(oracle, tw_tables, tmp_path):
    from ziren_b200 import keccak_sponge as ks
    from ziren_b200 import synthetic
    b = ks.synthetic_blocks(2, [1, 2], seed=3)
    t = oracle.keccak_sponge_trace(b, 1 << ks.padded_log_height(len(b)))
    case = synthetic.keccak_real_case(b, t, log_cpu=10, num_queries=6, pow_bits=3)
    om = oracle.OracleMachine(case.machine)
    om.setup(case.prep)
    _check_chip(oracle, om, case.machine, "KeccakSponge", None, case.traces["KeccakSponge"], case.public_values, tw_tables, tmp_path)
    # the extended Byte table: preprocessed columns, four receives in batches of two
    _check_chip(oracle, om, case.machine, "Byte", case.prep["Byte"], case.traces["Byte"], case.public_values, tw_tables, tmp_path)


@pytest.mark.parametrize("name", ["Cpu", "Global", "Fibonacci", "Sink"])
def test_generated_kernels_of_the_mini_machine_match_the_oracle_on_the_host(oracle, tw_tables, tmp_path, name):
    from ziren_b200 import synthetic
    case = synthetic.mini_case(seed=11)
    om = oracle.OracleMachine(case.machine)
    om.setup(case.prep)
    _check_chip(oracle, om, case.machine, name, case.prep.get(name), case.traces[name], case.public_values, tw_tables, tmp_path)


def test_generated_kernel_of_the_real_global_chip_matches_the_oracle_on_the_host(oracle, tw_tables, tmp_path):
    """GlobalChip::eval restated (ziren_b200/synthetic.py _global_chip: 120 constraints of degree three over septic-extension
    products, first-row and transition selectors, the global-scope cumulative sum) over rows from trace generation."""
    from ziren_b200 import synthetic
    from ziren_b200 import tracegen as tg
    ev = tg.synthetic_global_events(50, seed=4)
    case = synthetic.global_case(oracle.global_trace(ev, 64), num_queries=6, pow_bits=3)
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    _check_chip(oracle, om, case.machine, "Global", None, case.traces["Global"], case.public_values, tw_tables, tmp_path)


def test_generated_kernels_of_the_memory_global_tables_match_the_oracle_on_the_host(oracle, tw_tables, tmp_path):
    """MemoryGlobalChip::eval restated (320 constraints, public values, next-row columns, one send per row) and the Global chip
    with its receive: constraint kernel and permutation-trace kernel of each."""
    from ziren_b200 import synthetic
    from ziren_b200 import tracegen as tg
    init = tg.synthetic_memory_global_events(20, seed=1)
    init[:, 2], init[:, 3] = 0, 1
    fin = tg.synthetic_memory_global_events(30, seed=2)
    pi, pf = int(init[:, 0].min()) - 1, int(fin[:, 0].min()) - 1
    gev = np.concatenate([tg.memory_global_lookup_events(init, False), tg.memory_global_lookup_events(fin, True)])
    si, sf = init[np.argsort(init[:, 0])], fin[np.argsort(fin[:, 0])]
    case = synthetic.memory_global_case(oracle.memory_global_trace(si, pi, 32), oracle.memory_global_trace(sf, pf, 32),
                                        oracle.global_trace(gev, 64), pi, pf, num_queries=6, pow_bits=3)
    om = oracle.OracleMachine(case.machine)
    om.setup({})
    for name in ("MemoryGlobalInit", "MemoryGlobalFinalize", "Global"):
        _check_chip(oracle, om, case.machine, name, None, case.traces[name], case.public_values, tw_tables, tmp_path)


@pytest.mark.parametrize("which,name", [("edge", "PlainA"), ("edge", "Alu"), ("edge", "Program"), ("compress", "Poseidon2Wide"),
                                        ("compress", "PublicValues"), ("core", "MemoryInstrs"), ("core", "Byte")])
def test_generated_kernels_of_other_machines_match_the_oracle_on_the_host(oracle, tw_tables, tmp_path, which, name):
    """Chips without lookups (no LogUp code at all), preprocessed-only tables, the wide recursion-like and core-like tables."""
    from ziren_b200 import synthetic
    case = {"edge": lambda: synthetic.edge_case(), "compress": lambda: synthetic.compress_case(log_max=7),
            "core": lambda: synthetic.core_case(log_cpu=8)}[which]()
    om = oracle.OracleMachine(case.machine)
    om.setup(case.prep)
    _check_chip(oracle, om, case.machine, name, case.prep.get(name), case.traces[name], case.public_values, tw_tables, tmp_path)


def _random_chip(rng, name, degree_cap, lookups=True):
    """A chip with random constraints and lookups: templates of random expressions (degree <= degree_cap) stamped over
    several column offsets (so that the generator finds repeated shapes), a few one-off constraints, and an odd or even number
    of sends / receives whose values are constants, columns, and affine combinations."""
    from ziren_b200.air import KIND_BYTE, KIND_MEMORY, Chip
    prep_w, main_w = int(rng.integers(0, 4)), int(rng.integers(8, 14))
    n_templates, reps = int(rng.integers(2, 5)), int(rng.integers(3, 7))
    n_lookups = int(rng.integers(1, 10)) if lookups else 0
    seeds = rng.integers(0, 1 << 30, 64)

    def ev(b):
        def leaf(r, off):
            k = int(r.integers(0, 10))
            if k <= 4:
                return b.main(int(r.integers(0, main_w - 7)) + off, next=bool(r.integers(0, 4) == 0))
            if k == 5 and prep_w:
                return b.prep(int(r.integers(0, prep_w)), next=bool(r.integers(0, 2)))
            if k == 6:
                return b.const(int(r.integers(0, 1 << 31)))
            if k == 7:
                return b.pub(int(r.integers(0, 4)))
            if k == 8:
                return [b.is_first_row(), b.is_last_row(), b.is_transition()][int(r.integers(0, 3))]
            return b.main(int(r.integers(0, main_w - 7)) + off)

        def expr(r, off, depth):
            if depth == 0:
                return leaf(r, off)
            op = int(r.integers(0, 4))
            x = expr(r, off, depth - 1)
            if op == 3:
                return -x
            y = expr(r, off, depth - 1)
            if op == 2:
                if x.deg + y.deg > degree_cap:
                    return x + y
                return x * y
            return x + y if op == 0 else x - y
        for t in range(n_templates):
            for off in range(reps):                      # the same expression over shifted columns: one shape, `reps` members
                b.assert_zero(expr(np.random.default_rng(int(seeds[t])), off, 3))
        for t in range(3):                               # one-off constraints
            b.assert_zero(expr(np.random.default_rng(int(seeds[10 + t])), 0, 2))
        if degree_cap >= 4:                              # a product of degree_cap columns: more than two quotient chunks
            prod = b.main(0)
            for k in range(1, degree_cap):
                prod = prod * b.main(k % main_w, next=bool(k == 2))
            b.assert_zero(prod - b.main(main_w - 1))
        r = np.random.default_rng(int(seeds[20]))
        for i in range(n_lookups):
            vals = []
            for _ in range(int(r.integers(1, 5))):
                k = int(r.integers(0, 4))
                c0 = b.main(int(r.integers(0, main_w)))
                vals.append(int(r.integers(0, 300)) if k == 0 else c0 if k == 1 else c0 + int(r.integers(1, 99)) if k == 2
                            else c0 * int(r.integers(2, 9)) + b.main(int(r.integers(0, main_w))))
            mult = [1, b.main(int(r.integers(0, main_w))), b.main(int(r.integers(0, main_w))) * 3 + 2][int(r.integers(0, 3))]
            (b.send if i % 3 else b.receive)(KIND_MEMORY if i % 2 else KIND_BYTE, vals, mult)
    return Chip(name, prep_w, main_w, ev)


@pytest.mark.parametrize("seed,log_blowup,degree_cap,lookups", [(1, 1, 3, True), (2, 1, 3, True), (3, 2, 5, True), (4, 3, 5, True),
                                                                (5, 1, 2, False), (6, 2, 4, True), (7, 3, 9, True)])
def test_generated_kernels_of_random_chips_match_the_oracle_on_the_host(oracle, tw_tables, tmp_path, seed, log_blowup, degree_cap,
                                                                        lookups):
    """Generator fuzz: random constraint DAGs (repeated shapes and one-off constraints, every leaf kind, degrees up to 5 with
    quotient chunk counts 1, 2 and 4) and random lookups (batches of 1, 2 and 4; constants, columns and affine values; odd
    counts), evaluated on random traces: the generated K3 and K5 source must reproduce the oracle exactly."""
    from ziren_b200.air import Machine
    rng = np.random.default_rng(1000 + seed)
    chip = _random_chip(rng, "Fuzz", degree_cap, lookups)
    assert chip.log_quotient_degree == {2: 0 if not lookups else 1, 3: 1, 4: 2, 5: 2, 9: 3}[degree_cap]
    machine = Machine([chip], num_pv_elts=4, log_blowup=log_blowup, num_queries=4, pow_bits=2)
    n = 1 << int(rng.integers(3, 7))
    trace = kb.random_elements(rng, (n, chip.main_width))
    prep = kb.random_elements(rng, (n, chip.prep_width)) if chip.prep_width else None
    pv = kb.random_elements(rng, 8)
    om = oracle.OracleMachine(machine)
    _check_chip(oracle, om, machine, "Fuzz", prep, trace, pv, tw_tables, tmp_path, seed=seed)

"""CPU tests of the host side: chips-as-data builder, descriptor, ABI surface, sharding."""
import os
import re

import numpy as np
import pytest

from ziren_b200 import air, synthetic
from ziren_b200 import field as kb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_degree_rule_matches_chip_new():
    # crates/stark/src/chip.rs:72-87
    def deg2(b):
        b.assert_zero(b.main(0) * b.main(1) - b.main(2))
    def deg3(b):
        b.assert_zero(b.main(0) * b.main(1) * b.main(2) - b.main(0))
    def deg5(b):
        x = b.main(0)
        b.assert_zero(x * x * x * x * x - x)
    assert air.Chip("a", 0, 3, deg2).log_quotient_degree == 0
    assert air.Chip("b", 0, 3, deg3).log_quotient_degree == 1
    assert air.Chip("c", 0, 3, deg5).log_quotient_degree == 2
    def with_lookup(b):
        deg2(b)
        b.send(air.KIND_BYTE, [b.main(0)], 1)
    c = air.Chip("d", 0, 3, with_lookup)
    assert c.log_quotient_degree == 1 and c.perm_width_ef == 2 and c.num_constraints == 1 + 1 + 3


def test_permutation_width_and_constraint_count():
    # crates/stark/src/permutation.rs:18-23 and :355-389
    case = synthetic.mini_case()
    cpu = case.machine.chip("Cpu")
    assert cpu.num_local_lookups == 4 and cpu.perm_width_ef == 3
    glob = case.machine.chip("Global")
    assert glob.num_constraints == len(glob.builder.constraints) + (glob.perm_width_ef - 1 + 3) + 14
    # global-scope lookups do not enter the local permutation
    assert glob.num_local_lookups == 1 and len(glob.builder.sends) == 2


def test_lookups_must_be_affine():
    def bad(b):
        b.send(air.KIND_BYTE, [b.main(0) * b.main(1)], 1)
    with pytest.raises(ValueError):
        air.Chip("x", 0, 2, bad)


def test_descriptor_roundtrips_through_the_oracle_parser(oracle):
    case = synthetic.mini_case()
    om = oracle.OracleMachine(case.machine)
    for c in case.machine.chips:
        info = om.chip_info(c.name)
        assert info["perm_width_ef"] == c.perm_width_ef
        assert info["num_constraints"] == c.num_constraints
        assert info["log_quotient_degree"] == c.log_quotient_degree


def test_synthetic_traces_satisfy_their_airs():
    case = synthetic.mini_case(seed=5)
    t = case.traces["Cpu"].astype(np.uint64)
    for g in range(3):
        a, b, c, d, e = (t[:, 6 * g + k] for k in range(5))
        assert np.array_equal(a * b % kb.P, c) and np.array_equal(c * d % kb.P, e)
    assert np.array_equal(t[1:, 0], t[:-1, 0] + 1)
    sent = sum(np.bincount(case.traces[n][:, 6 * g + 5], minlength=1 << 16)
               for n, ng in (("Cpu", 3), ("AddSub", 1), ("Global", 1)) for g in range(ng))
    assert np.array_equal(sent, case.traces["Byte"][:, 0])
    fib = case.traces["Fibonacci"].astype(np.uint64)
    assert np.array_equal((fib[:-1, 0] + fib[:-1, 1]) % kb.P, fib[1:, 1])


def test_reference_shapes_match_costs():
    case = synthetic.fibonacci_core_case(log_cpu=8)
    assert case.machine.chip("Cpu").cost == 119           # mips_costs.json
    assert case.machine.chip("AddSub").cost == 47
    big = synthetic.tune_wide("KeccakSponge", 10, 4259, 40)
    assert 6 * big.groups + big.extra + 4 * 21 + 8 == 4259


def test_montgomery_roundtrip():
    rng = np.random.default_rng(0)
    x = kb.random_elements(rng, 1000)
    assert np.array_equal(kb.from_monty(kb.to_monty(x)), x)
    assert kb.to_monty(np.array([1], np.uint32))[0] == 0x1FFFFFE   # kb31_t.hpp:29


def test_library_exports_every_declared_symbol():
    """The C-ABI library loads on a machine without a GPU and exports every function that
    include/zkb200.h declares (no compute calls here)."""
    from ziren_b200 import _ffi
    header = open(os.path.join(ROOT, "include", "zkb200.h")).read()
    declared = set(re.findall(r"\b(zkb200_[a-z0-9_]+)\s*\(", header))
    declared -= {"zkb200_ctx", "zkb200_pk", "zkb200_shard", "zkb200_trace"}
    assert declared == set(_ffi.SIGNATURES), declared ^ set(_ffi.SIGNATURES)
    lib = _ffi.lib()
    for name in declared:
        assert hasattr(lib, name), name


def test_no_product_code_touches_the_oracle():
    """The product path must not import, link or call anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ziren_b200")):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.replace("oracle/", "oracle/").lower() or f in ("synthetic.py",) or \
                    all("import" not in line and "#include" not in line and "CDLL" not in line
                        for line in text.splitlines() if "oracle" in line.lower()), f


def test_ctx_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this checks the no-GPU behaviour")
    from ziren_b200.prover import B200Prover, ZkbError
    with pytest.raises(ZkbError, match="CUDA"):
        B200Prover(synthetic.mini_case().machine)


def test_shard_assignment_is_a_partition():
    from ziren_b200.sharding import assign_shards
    for n in (1, 5, 18):
        for world in (1, 2, 4, 8):
            got = sorted(sum((assign_shards(n, world, r) for r in range(world)), []))
            assert got == list(range(n))


def _gather_worker(rank, world, port, q):
    import torch.distributed as dist
    from ziren_b200.sharding import assign_shards, gather_commitments
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n = 5
    mine = {i: np.full(8, 100 * i + 7, np.uint32) + np.arange(8, dtype=np.uint32) for i in assign_shards(n, world, rank)}
    table = gather_commitments(mine, n)
    q.put((rank, table.tolist()))
    dist.destroy_process_group()


def test_commitment_gather_world_size_2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    want = [[100 * i + 7 + k for k in range(8)] for i in range(5)]
    for _, table in res:
        assert table == want


def test_lane_pool_hands_out_any_free_lane(host):
    """csrc/lane_pool.h under contention: one holder per lane, never more than `active` lanes held, and the
    calls return (a waiter is woken by ANY release, not only lane 0's)."""
    for threads, active in ((4, 3), (8, 3), (2, 1), (6, 4), (3, 3)):
        assert host.hostcheck_lane_pool(threads, 3000, active) == 0, (threads, active)



def test_constraint_kernels_are_generated_and_compile_without_a_gpu():
    """K3 code generation (csrc/quotient_codegen.cpp): every chip of a machine becomes CUDA source that NVRTC
    compiles to an sm_100a cubin; needs no device."""
    import ctypes as C
    from ziren_b200 import _ffi, synthetic
    m = synthetic.mini_case().machine
    desc = np.ascontiguousarray(m.descriptor(), dtype=np.uint32)
    n = C.c_size_t()
    rc = _ffi.lib().zkb200_codegen_compile_check(desc.ctypes.data_as(_ffi.u32p), desc.size, C.byref(n))
    assert rc == len(m.chips), _ffi.lib().zkb200_last_error(None).decode()
    assert n.value > 10000


def test_zkpf_decoder_follows_the_reference_proof_types(oracle):
    """ziren_b200/proof.py (the Python twin of shim/src/proof.rs::decode_zkpf) yields the fields of the reference's
    ShardProof / ChipOpenedValues / AirOpenedValues (crates/stark/src/types.rs:38-83) and of Plonky3's FriProof
    (mirrored in crates/recursion/circuit/src/types.rs:35-75), in that nesting; when the reference checkout is
    present the field names are read from it."""
    from ziren_b200 import proof as zkproof
    case = synthetic.mini_case()
    om = oracle.OracleMachine(case.machine)
    om.setup(case.prep)
    words, _ = om.prove_shard(case.traces, case.public_values)
    p = zkproof.parse(words)
    want = {"ShardProof": ["commitment", "opened_values", "opening_proof", "chip_ordering", "public_values"],
            "ShardCommitment": ["main_commit", "permutation_commit", "quotient_commit"],
            "ChipOpenedValues": ["preprocessed", "main", "permutation", "quotient", "global_cumulative_sum",
                                 "local_cumulative_sum", "log_degree"],
            "AirOpenedValues": ["local", "next"]}
    ref = "/root/reference/crates/stark/src/types.rs"
    if os.path.exists(ref):
        src = open(ref).read()
        for struct, fields in want.items():
            body = re.search(r"pub struct %s<[^{]*\{(.*?)\n\}" % struct, src, re.S).group(1)
            assert re.findall(r"pub (\w+):", body) == fields, struct
    assert list(p) == want["ShardProof"]
    assert list(p["commitment"]) == want["ShardCommitment"]
    chip = p["opened_values"]["chips"][0]
    assert want["ChipOpenedValues"] == [k for k in chip if k != "name"]
    assert list(chip["main"]) == want["AirOpenedValues"]
    fri = p["opening_proof"]
    assert set(fri) == {"commit_phase_commits", "query_proofs", "final_poly", "pow_witness"}
    q = fri["query_proofs"][0]
    assert list(q) == ["input_proof", "commit_phase_openings"]
    assert list(q["input_proof"][0]) == ["opened_values", "opening_proof"]
    assert list(q["commit_phase_openings"][0]) == ["sibling_value", "opening_proof"]
    # the Rust decoder reads the words in the same order: its source mentions every field
    rust = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "shim", "src", "proof.rs")).read()
    for name in sum(want.values(), []) + ["commit_phase_commits", "query_proofs", "final_poly", "pow_witness", "sibling_value"]:
        assert name in rust, name


def test_constraint_generator_groups_the_keccak_chip_by_shape(tmp_path, monkeypatch):
    """K3 generator (csrc/quotient_codegen.cpp): the real KeccakSponge chip's 3 788 constraints become a few loops over
    parameter tables (one per expression shape) instead of straight-line code per constraint; the kernel compiles with
    NVRTC for sm_100a without a GPU."""
    import ctypes as C
    from ziren_b200 import _ffi, keccak_sponge, synthetic
    monkeypatch.setenv("ZKB200_CODEGEN_DUMP", str(tmp_path))
    m = synthetic.keccak_real_case(keccak_sponge.synthetic_blocks(1, 1), None, log_cpu=10).machine
    desc = np.ascontiguousarray(m.descriptor(), dtype=np.uint32)
    n = C.c_size_t()
    assert _ffi.lib().zkb200_codegen_compile_check(desc.ctypes.data_as(_ffi.u32p), desc.size, C.byref(n)) == len(m.chips)
    src = (tmp_path / "qk_KeccakSponge.cu").read_text()
    n_shapes = src.count("static __device__ __noinline__ Ef shape")
    assert 10 <= n_shapes <= 40 and src.count("static __device__ __noinline__ Ef seg") <= 3
    assert "QK_TAB[" in src and src.count("\n") < 12000          # the per-constraint version was ~70 k lines

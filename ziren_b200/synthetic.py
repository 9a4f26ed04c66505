"""Synthetic machines and shard traces with the reference's shapes.

No guest ELF can be built in this environment (no MIPS/Rust toolchain), so benchmark and test
shards are synthetic: satisfiable AIRs whose table names, heights and per-row column budgets
follow the reference (crates/core/executor/src/artifacts/mips_costs.json — cost = P + M + 4E + 4Q
per crates/stark/src/chip.rs:154-163 — and crates/core/machine/src/shape/maximal_shapes.json),
filled from a seeded RNG.  Structure mirrors the real machine: a preprocessed byte/range table
and a preprocessed program table that receive lookups from the execution tables, one
global-scope table whose last 14 columns carry the septic digest, public values.

Small real AIR: Fibonacci with public values, crates/stark/src/stark_testing.rs:25-61.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import field as kb
from .air import (KIND_BYTE, KIND_GLOBAL, KIND_MEMORY, KIND_PROGRAM, KIND_SYSCALL, KIND_SYSCALL_RESULT, SCOPE_GLOBAL, Chip, Machine)

P = kb.P
RANGE_BITS = 16  # the reference's Byte table has 2^16 rows


@dataclass
class WideSpec:
    """A 'wide' execution table: `groups` column groups (a,b,c,d,e,r) with a*b=c, c*d=e,
    the first `lookups` groups send r to the byte table, `extra` unconstrained columns.
    `counter`: group 0's `a` starts at public value 0 and increments (transition constraint) and
    is looked up in the program table.  `global_scope`: 14 trailing digest columns."""
    name: str
    log_height: int
    groups: int
    lookups: int
    extra: int = 0
    counter: bool = False
    global_scope: bool = False

    @property
    def width(self) -> int:
        return 6 * self.groups + self.extra + (14 if self.global_scope else 0)


# Constraints per 6-column group of the wide tables: 2 (default, "sparse": a*b = c, c*d = e) up to 12
# ("dense", about 2 per column as SURVEY.md section 8d(i) asks for a K3 throughput figure: the real
# KeccakSponge chip has 1 388 constraints over 4 167 columns, the ALU chips are denser still).  The extra
# constraints are further degree <= 3 consequences of the same two identities, so the same traces satisfy them.
CONSTRAINTS_PER_GROUP = 2


def _wide_chip(s: WideSpec) -> Chip:
    def ev(b):
        for g in range(s.groups):
            a, bb, c, d, e, r = (b.main(6 * g + k) for k in range(6))
            x, y = a * bb - c, c * d - e
            extra = [lambda: x * d, lambda: y * a, lambda: a * bb * d - e, lambda: x * r, lambda: y * r, lambda: x + y,
                     lambda: x * a, lambda: y * c, lambda: x * e, lambda: y * bb]
            b.assert_zero(x)
            b.assert_zero(y)
            for k in range(max(0, min(CONSTRAINTS_PER_GROUP, 12) - 2)):
                b.assert_zero(extra[k]())
            if g < s.lookups:
                b.send(KIND_BYTE, [r], 1)
        if s.counter:
            a0 = b.main(0)
            b.when_first_row().assert_eq(a0, b.pub(0))
            b.when_transition().assert_eq(b.main(0, next=True), a0 + 1)
            b.send(KIND_PROGRAM, [a0], 1)
        if s.global_scope:
            b.send(KIND_GLOBAL, [b.main(1)], 1, scope=SCOPE_GLOBAL)
    return Chip(s.name, 0, s.width, ev, global_scope=s.global_scope)


def _wide_trace(s: WideSpec, rng: np.random.Generator, pv0: int) -> np.ndarray:
    n = 1 << s.log_height
    t = np.empty((n, s.width), dtype=np.uint32)
    for g in range(s.groups):
        a = kb.random_elements(rng, n)
        if g == 0 and s.counter:
            a = ((np.arange(n, dtype=np.uint64) + np.uint64(pv0)) % np.uint64(P)).astype(np.uint32)
        bb, d = kb.random_elements(rng, n), kb.random_elements(rng, n)
        c = kb.mul(a, bb)
        t[:, 6 * g + 0], t[:, 6 * g + 1], t[:, 6 * g + 2] = a, bb, c
        t[:, 6 * g + 3], t[:, 6 * g + 4] = d, kb.mul(c, d)
        t[:, 6 * g + 5] = rng.integers(0, 1 << RANGE_BITS, size=n, dtype=np.uint32)
    rest = s.width - 6 * s.groups
    if rest:
        t[:, 6 * s.groups:] = kb.random_elements(rng, (n, rest))
    return t


def _byte_chip() -> Chip:
    def ev(b):
        b.receive(KIND_BYTE, [b.prep(0)], b.main(0))
    return Chip("Byte", 1, 1, ev, local_only=True)


def _program_chip() -> Chip:
    def ev(b):
        b.receive(KIND_PROGRAM, [b.prep(0)], b.main(0))
    return Chip("Program", 2, 1, ev, local_only=True)


def _fib_chip() -> Chip:
    # stark_testing.rs:35-61, public values at offsets 1..3, plus a send of every row
    def ev(b):
        left, right = b.main(0), b.main(1)
        nl, nr = b.main(0, next=True), b.main(1, next=True)
        b.when_first_row().assert_eq(left, b.pub(1))
        b.when_first_row().assert_eq(right, b.pub(2))
        b.when_transition().assert_eq(right, nl)
        b.when_transition().assert_eq(left + right, nr)
        b.when_last_row().assert_eq(right, b.pub(3))
        b.send(KIND_MEMORY, [left, right], 1)
    return Chip("Fibonacci", 0, 2, ev)


def _sink_chip() -> Chip:
    def ev(b):
        b.receive(KIND_MEMORY, [b.main(0), b.main(1)], b.main(2))
    return Chip("Sink", 0, 3, ev, local_only=True)


class ShardCase:
    """A machine plus one satisfiable shard: preprocessed traces, main traces, public values.
    Traces are row-major uint32 arrays in CANONICAL form; use `.monty()` for the C ABI."""

    def __init__(self, machine: Machine, prep: dict, traces: dict, public_values: np.ndarray, cycles: int):
        self.machine, self.prep, self.traces = machine, prep, traces
        self.public_values, self.cycles = public_values, cycles

    @property
    def cells(self) -> int:
        tot = 0
        for name, t in self.traces.items():
            tot += t.shape[0] * self.machine.chip(name).cost
        return tot

    @property
    def trace_bytes(self) -> int:
        return sum(4 * t.size for t in self.traces.values())


def _plain_chip(name: str) -> Chip:
    """No lookups at all and degree 2: log_quotient_degree 0, zero-width permutation trace."""
    def ev(b):
        x, y, z = b.main(0), b.main(1), b.main(2)
        b.assert_zero(x * y - z)
        b.when_transition().assert_eq(b.main(0, next=True), x + 2)
    return Chip(name, 0, 3, ev)


def _plain_trace(rng, log_height: int) -> np.ndarray:
    n = 1 << log_height
    t = np.empty((n, 3), dtype=np.uint32)
    t[:, 0] = ((np.arange(n, dtype=np.uint64) * 2 + 5) % P).astype(np.uint32)
    t[:, 1] = kb.random_elements(rng, n)
    t[:, 2] = kb.mul(t[:, 0], t[:, 1])
    return t


def build_case(wides: list[WideSpec], *, with_fib: int | None = None, seed: int = 0xC0FFEE,
               num_queries: int = 84, pow_bits: int = 16, cycles: int | None = None, log_blowup: int = 1,
               plain: list[tuple[str, int]] = (), real: dict | None = None) -> ShardCase:
    """Assemble a machine from wide tables (+ optional Fibonacci/Sink pair of log height
    `with_fib`) and generate one balanced shard.  `real`: restated real chips next to the synthetic ones
    (keccak_real_case): {"chips": [...], "traces": {...}, "byte_mults": (65536, 3)}; the Byte table is then the one of
    ziren_b200/keccak_air.py, which also receives their XOR / range lookups."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pv = np.zeros(8, dtype=np.uint32)
    chips, traces, prep = [], {}, {}
    counter_specs = [w for w in wides if w.counter]
    assert len(counter_specs) <= 1
    byte_counts = np.zeros(1 << RANGE_BITS, dtype=np.uint64)
    for w in wides:
        chips.append(_wide_chip(w))
        t = _wide_trace(w, rng, int(pv[0]))
        if w.global_scope:
            t[-1, -14:] = kb.random_elements(rng, 14)
        traces[w.name] = t
        for g in range(w.lookups):
            byte_counts += np.bincount(t[:, 6 * g + 5], minlength=1 << RANGE_BITS).astype(np.uint64)
    if real is None:
        chips.append(_byte_chip())
        prep["Byte"] = np.arange(1 << RANGE_BITS, dtype=np.uint32).reshape(-1, 1)
        traces["Byte"] = (byte_counts % np.uint64(P)).astype(np.uint32).reshape(-1, 1)
    else:
        from . import keccak_air
        chips.append(keccak_air.byte_chip())
        prep["Byte"] = keccak_air.byte_prep()
        mults = np.concatenate([byte_counts.reshape(-1, 1), np.asarray(real["byte_mults"], dtype=np.uint64)], axis=1)
        traces["Byte"] = (mults % np.uint64(P)).astype(np.uint32)
        chips.extend(real["chips"])
        traces.update(real["traces"])
    if counter_specs:
        n = 1 << counter_specs[0].log_height
        chips.append(_program_chip())
        pt = np.empty((n, 2), dtype=np.uint32)
        pt[:, 0] = np.arange(n, dtype=np.uint32)
        pt[:, 1] = kb.random_elements(rng, n)
        prep["Program"] = pt
        traces["Program"] = np.ones((n, 1), dtype=np.uint32)
    if with_fib is not None:
        n = 1 << with_fib
        a, b_ = 0, 1
        rows = np.empty((n, 2), dtype=np.uint32)
        for i in range(n):
            rows[i] = (a, b_)
            a, b_ = b_, (a + b_) % P
        pv[1], pv[2], pv[3] = 0, 1, rows[-1, 1]
        chips.append(_fib_chip())
        traces["Fibonacci"] = rows
        chips.append(_sink_chip())
        sink = np.zeros((2 * n, 3), dtype=np.uint32)
        perm = rng.permutation(n)
        sink[:n, :2] = rows[perm]
        sink[:n, 2] = 1
        traces["Sink"] = sink
    for name, lh in plain:
        chips.append(_plain_chip(name))
        traces[name] = _plain_trace(rng, lh)
    machine = Machine(chips, num_pv_elts=4, num_queries=num_queries, pow_bits=pow_bits, log_blowup=log_blowup)
    if cycles is None:
        cycles = (1 << counter_specs[0].log_height) if counter_specs else 0
    return ShardCase(machine, prep, traces, pv, cycles)


def tune_wide(name: str, log_height: int, cost: int, lookups: int, **kw) -> WideSpec:
    """Pick (groups, extra) so that the table's per-row cost P+M+4E+4Q matches `cost`
    (mips_costs.json) for the given number of byte lookups."""
    n_lk = lookups + (1 if kw.get("counter") else 0)
    e = -(-n_lk // 2) + 1
    m = cost - 4 * e - 8 - (14 if kw.get("global_scope") else 0)
    groups = max(lookups, m // 6, 1)
    extra = max(0, m - 6 * groups)
    return WideSpec(name, log_height, groups, lookups, extra, **kw)


# -- named workloads ---------------------------------------------------------------------------
def mini_case(seed: int = 1, **kw) -> ShardCase:
    """Tiny machine covering every structural feature (used by parity tests)."""
    wides = [WideSpec("Cpu", 6, 3, 3, 2, counter=True), WideSpec("AddSub", 5, 2, 1, 0),
             WideSpec("Global", 4, 1, 1, 1, global_scope=True)]
    kw.setdefault("num_queries", 8)
    kw.setdefault("pow_bits", 4)
    return build_case(wides, with_fib=4, seed=seed, **kw)


def edge_case(seed: int = 7, **kw) -> ShardCase:
    """Edge shapes: chips with no lookups (log_quotient_degree 0, zero-width permutation matrices,
    one of them alone at its height), tiny tables (2^1 rows), equal heights ordered by name."""
    wides = [WideSpec("Cpu", 5, 2, 2, 1, counter=True), WideSpec("Alu", 5, 1, 1, 0)]
    kw.setdefault("num_queries", 6)
    kw.setdefault("pow_bits", 3)
    return build_case(wides, seed=seed, plain=[("PlainA", 7), ("PlainB", 5), ("PlainTiny", 1)], **kw)


def noprep_case(seed: int = 5, **kw) -> ShardCase:
    """A machine with NO preprocessed table at all (empty proving-key commitment, three opening
    rounds instead of four): Fibonacci + Sink + two plain chips."""
    rng = np.random.Generator(np.random.PCG64(seed))
    kw.setdefault("num_queries", 5)
    kw.setdefault("pow_bits", 3)
    n = 1 << 5
    a, b_ = 0, 1
    rows = np.empty((n, 2), dtype=np.uint32)
    for i in range(n):
        rows[i] = (a, b_)
        a, b_ = b_, (a + b_) % P
    pv = np.zeros(8, dtype=np.uint32)
    pv[1], pv[2], pv[3] = 0, 1, rows[-1, 1]
    sink = np.zeros((n, 3), dtype=np.uint32)
    sink[:, :2] = rows[rng.permutation(n)]
    sink[:, 2] = 1
    chips = [_fib_chip(), _sink_chip(), _plain_chip("PlainA"), _plain_chip("PlainB")]
    traces = {"Fibonacci": rows, "Sink": sink, "PlainA": _plain_trace(rng, 6), "PlainB": _plain_trace(rng, 3)}
    machine = Machine(chips, num_pv_elts=4, num_queries=kw["num_queries"], pow_bits=kw["pow_bits"],
                      log_blowup=kw.get("log_blowup", 1))
    return ShardCase(machine, {}, traces, pv, 0)


def fibonacci_core_case(log_cpu: int = 16, seed: int = 0xC0FFEE, **kw) -> ShardCase:
    """S1 'fib-2^16' (SURVEY.md §8d): a single small core shard."""
    h = log_cpu
    wides = [tune_wide("Cpu", h, 119, 10, counter=True), tune_wide("AddSub", h - 1, 47, 4),
             tune_wide("Lt", h - 2, 52, 4), tune_wide("Branch", h - 3, 90, 6),
             tune_wide("MemoryInstrs", h - 4, 115, 8), tune_wide("Global", h - 4, 115, 6, global_scope=True),
             tune_wide("MemoryLocal", h - 6, 100, 6)]
    return build_case(wides, seed=seed, **kw)


# maximal_shapes.json["21"][1]
_CORE21 = {"Cpu": 21, "AddSub": 20, "MemoryInstrs": 19, "Bitwise": 19, "Lt": 18, "ShiftRight": 18, "Global": 18,
           "Branch": 17, "ShiftLeft": 16, "MemoryLocal": 15, "MovCond": 15, "Jump": 15, "Mul": 13,
           "SyscallInstrs": 12, "SyscallCore": 12, "MiscInstrs": 8, "CloClz": 8, "DivRem": 4}
_COSTS = {"Cpu": 119, "AddSub": 47, "MemoryInstrs": 115, "Bitwise": 42, "Lt": 52, "ShiftRight": 131, "Global": 115,
          "Branch": 90, "ShiftLeft": 68, "MemoryLocal": 100, "MovCond": 48, "Jump": 82, "Mul": 110,
          "SyscallInstrs": 97, "SyscallCore": 39, "MiscInstrs": 152, "CloClz": 41, "DivRem": 162,
          "KeccakSponge": 4259}
_LOOKUPS = {"Cpu": 10, "AddSub": 4, "MemoryInstrs": 8, "Bitwise": 4, "Lt": 4, "ShiftRight": 8, "Global": 6,
            "Branch": 6, "ShiftLeft": 6, "MemoryLocal": 6, "MovCond": 4, "Jump": 6, "Mul": 8, "SyscallInstrs": 6,
            "SyscallCore": 3, "MiscInstrs": 8, "CloClz": 3, "DivRem": 10, "KeccakSponge": 40}


def core_case(log_cpu: int = 21, seed: int = 0xC0FFEE, **kw) -> ShardCase:
    """S3 'core-2^21' tendermint-like maximal shard, scaled by log_cpu (heights shift together)."""
    d = 21 - log_cpu
    wides = []
    for name, lh in _CORE21.items():
        lh = max(4, lh - d)
        wides.append(tune_wide(name, lh, _COSTS[name], _LOOKUPS[name], counter=(name == "Cpu"),
                               global_scope=(name == "Global")))
    return build_case(wides, seed=seed, **kw)


def keccak_case(log_cpu: int = 20, log_keccak: int | None = None, seed: int = 0xC0FFEE, **kw) -> ShardCase:
    """S2 'keccak-2^20': a core shard whose area is dominated by the KeccakSponge precompile table
    (4259 columns per row).  Default keccak height = cpu height - 2."""
    lk = log_cpu - 2 if log_keccak is None else log_keccak
    h = log_cpu
    wides = [tune_wide("Cpu", h, 119, 10, counter=True), tune_wide("AddSub", h - 1, 47, 4),
             tune_wide("MemoryInstrs", h - 2, 115, 8), tune_wide("Bitwise", h - 2, 42, 4),
             tune_wide("Global", h - 3, 115, 6, global_scope=True), tune_wide("MemoryLocal", h - 5, 100, 6),
             tune_wide("SyscallInstrs", h - 8, 97, 6), tune_wide("SyscallCore", h - 8, 39, 3),
             tune_wide("KeccakSponge", lk, 4259, 40)]
    return build_case(wides, seed=seed, **kw)


def keccak_real_case(blocks: np.ndarray, keccak_trace: np.ndarray | None, log_cpu: int = 20, seed: int = 0xC0FFEE, **kw) -> ShardCase:
    """S2 'keccak-2^20' with the REAL KeccakSponge chip: the reference's 3531-column layout and its restated
    `Air::eval` (ziren_b200/keccak_air.py: 358 lookups, degree 3) instead of the synthetic 4167-column stand-in of
    keccak_case, next to the same synthetic core tables.  `blocks`: (n, 384) block records
    (ziren_b200/keccak_sponge.py); `keccak_trace`: the chip's rows in canonical form as a row filler wrote them
    (oracle in the tests, the CUDA kernel in bench.py), or None to leave the table out of `traces` (it is then handed to
    the prover as event records).  The Byte multiplicities, one MemoryLocalPrecompile row per memory access and one
    SyscallPrecompile row per event balance the chip's local lookups."""
    from . import keccak_air
    from . import keccak_sponge as ksp
    h = log_cpu
    wides = [tune_wide("Cpu", h, 119, 10, counter=True), tune_wide("AddSub", h - 1, 47, 4),
             tune_wide("MemoryInstrs", h - 2, 115, 8), tune_wide("Bitwise", h - 2, 42, 4),
             tune_wide("Global", h - 3, 115, 6, global_scope=True), tune_wide("MemoryLocal", h - 5, 100, 6),
             tune_wide("SyscallInstrs", h - 8, 97, 6), tune_wide("SyscallCore", h - 8, 39, 3)]
    mults, mem, sysc = keccak_air.receiver_tables(blocks)
    real_traces = {"MemoryLocalPrecompile": keccak_air.pad_rows(mem), "SyscallPrecompile": keccak_air.pad_rows(sysc)}
    if keccak_trace is not None:
        assert keccak_trace.shape[1] == ksp.WIDTH and keccak_trace.shape[0] >= len(blocks) * ksp.NUM_ROUNDS
        real_traces["KeccakSponge"] = keccak_trace
    real = {"chips": [keccak_air.keccak_sponge_chip(), keccak_air.memory_local_chip(), keccak_air.syscall_precompile_chip()],
            "traces": real_traces, "byte_mults": mults}
    return build_case(wides, seed=seed, real=real, **kw)


# recursion "compress" machine (BASELINE.json configs[4]); heights = shrink_shape of
# crates/recursion/core/src/machine.rs:155-172, main widths: Poseidon2Wide 313 and BatchFRI 13 from
# the reference (poseidon2_wide/columns/permutation.rs:22-34, chips/batch_fri.rs:39-56), the others
# are stand-ins of plausible size (the real widths need the Rust derive macros).
_COMPRESS = {"MemoryVar": (18, 8, 2), "Select": (18, 12, 2), "MemoryConst": (17, 6, 1), "BatchFRI": (17, 13, 2),
             "BaseAlu": (17, 12, 2), "ExpReverseBitsLen": (17, 12, 2), "Poseidon2Wide": (16, 313, 8), "ExtAlu": (15, 24, 3),
             "PublicValues": (4, 30, 1)}


def compress_case(log_max: int = 18, seed: int = 0xC0FFEE, **kw) -> ShardCase:
    """S5 'compress': one inner proof of the recursion machine, scaled so that the tallest table
    has 2^log_max rows."""
    d = 18 - log_max
    wides = []
    for name, (lh, width, lookups) in _COMPRESS.items():
        groups = max(lookups, width // 6, 1)
        wides.append(WideSpec(name, max(2, lh - d), groups, lookups, max(0, width - 6 * groups)))
    return build_case(wides, seed=seed, cycles=0, **kw)


# ---- real ALU chips (trace generation) ------------------------------------
# The arithmetic constraints of two reference chips restated in the AirBuilder, so that traces made by
# `zkb200_generate_alu_trace` can be proved and verified.  Their lookups (receive_instruction, byte
# range checks) need the Cpu and Byte tables of the full MIPS machine and are left out: these chips
# carry the chips' polynomial identities only.

def _assert_bool(b, x):
    b.assert_zero(x * (x - 1))


def _add_sub_chip() -> Chip:
    """AddSubChip::eval crates/core/machine/src/alu/add_sub/mod.rs:196-250 with AddOperation::eval
    crates/core/machine/src/operations/add.rs:57-94; columns of AddSubCols (add_sub/mod.rs:43-65)."""
    def ev(b):
        value = [b.main(2 + i) for i in range(4)]
        carry = [b.main(6 + i) for i in range(3)]
        x = [b.main(9 + i) for i in range(4)]
        y = [b.main(13 + i) for i in range(4)]
        is_add, is_sub = b.main(17), b.main(18)
        is_real = is_add + is_sub
        real = b.when(is_real)
        over = [x[0] + y[0] - value[0]] + [x[i] + y[i] - value[i] + carry[i - 1] for i in (1, 2, 3)]
        real.assert_zero(over[3] * (over[3] - 256))
        for i in range(3):
            real.assert_zero(carry[i] * (over[i] - 256))
        for i in range(3):
            real.assert_zero((carry[i] - 1) * over[i])
        for i in range(3):
            real.assert_zero(carry[i] * (carry[i] - 1))
        real.assert_zero(is_real * (is_real - 1))
        _assert_bool(b, is_add)
        _assert_bool(b, is_sub)
        _assert_bool(b, is_real)
    return Chip("AddSub", 0, 19, ev, local_only=True)


def _shift_left_chip() -> Chip:
    """ShiftLeft::eval crates/core/machine/src/alu/sll/mod.rs:285-410; columns of ShiftLeftCols."""
    def ev(b):
        a = [b.main(2 + i) for i in range(4)]
        bb = [b.main(6 + i) for i in range(4)]
        c = [b.main(10 + i) for i in range(4)]
        c_bits = [b.main(14 + i) for i in range(8)]
        by_bits = [b.main(22 + i) for i in range(8)]
        mult = b.main(30)
        res = [b.main(31 + i) for i in range(4)]
        car = [b.main(35 + i) for i in range(4)]
        by_bytes = [b.main(39 + i) for i in range(4)]
        total = c_bits[0]
        for i in range(1, 8):
            total = total + c_bits[i] * (1 << i)
        b.assert_eq(total, c[0])
        nbits = c_bits[0] + c_bits[1] * 2 + c_bits[2] * 4
        for i in range(8):
            b.when(by_bits[i]).assert_eq(nbits, i)
        for i in range(8):
            b.when(by_bits[i]).assert_eq(mult, 1 << i)
        for i in range(4):
            v = bb[i] * mult - car[i] * 256
            if i > 0:
                v = v + car[i - 1]
            b.assert_eq(res[i], v)
        nbytes = c_bits[3] + c_bits[4] * 2
        for i in range(4):
            b.when(by_bytes[i]).assert_eq(nbytes, i)
        for k in range(4):
            sh = b.when(by_bytes[k])
            for i in range(4):
                if i < k:
                    sh.assert_eq(a[i], 0)
                else:
                    sh.assert_eq(a[i], res[i - k])
        for bit in c_bits:
            _assert_bool(b, bit)
        for s in by_bits:
            _assert_bool(b, s)
        tot = by_bits[0]
        for s in by_bits[1:]:
            tot = tot + s
        b.assert_eq(tot, 1)
        for s in by_bytes:
            _assert_bool(b, s)
        tot = by_bytes[0]
        for s in by_bytes[1:]:
            tot = tot + s
        b.assert_eq(tot, 1)
        _assert_bool(b, b.main(43))
    return Chip("ShiftLeft", 0, 44, ev, local_only=True)


def _lt_chip() -> Chip:
    """LtChip::eval crates/core/machine/src/alu/lt/mod.rs:288-470 without its byte lookups (b_masked / c_masked
    AND 0x7f, sltu = LTU of the comparison bytes) and its instruction receive; columns of LtCols."""
    def ev(b):
        is_slt, is_sltu = b.main(2), b.main(3)
        a = [b.main(4 + i) for i in range(4)]
        bw = [b.main(8 + i) for i in range(4)]
        cw = [b.main(12 + i) for i in range(4)]
        flags = [b.main(16 + i) for i in range(4)]
        b_masked, c_masked, not_eq_inv = b.main(20), b.main(21), b.main(22)
        msb_b, msb_c, bit_b, bit_c = b.main(23), b.main(24), b.main(25), b.main(26)
        sltu, is_comp_eq, is_sign_eq = b.main(27), b.main(28), b.main(29)
        cmp_b, cmp_c = b.main(30), b.main(31)
        is_real = is_slt + is_sltu
        b_comp = bw[:3] + [bw[3] * is_sltu + b_masked * is_slt]
        c_comp = cw[:3] + [cw[3] * is_sltu + c_masked * is_slt]
        b.assert_eq(bit_b, msb_b * is_slt)
        b.assert_eq(bit_c, msb_c * is_slt)
        inv_128 = pow(128, P - 2, P)
        b.assert_eq(msb_b, (bw[3] - b_masked) * inv_128)
        b.assert_eq(msb_c, (cw[3] - c_masked) * inv_128)
        _assert_bool(b, is_sign_eq)
        b.when(is_sign_eq).assert_eq(bit_b, bit_c)
        b.when(is_real).when(1 - is_sign_eq).assert_eq(bit_b + bit_c, 1)
        b.assert_eq(a[0], bit_b * (1 - bit_c) + is_sign_eq * sltu)
        for i in (1, 2, 3):
            b.assert_zero(a[i])
        sum_flags = flags[0] + flags[1] + flags[2] + flags[3]
        for f in flags:
            _assert_bool(b, f)
        _assert_bool(b, sum_flags)
        b.when(is_real).assert_eq(1 - is_comp_eq, sum_flags)
        _assert_bool(b, is_comp_eq)
        visited = None
        sel_b = sel_c = None
        for i in (3, 2, 1, 0):
            visited = flags[i] if visited is None else visited + flags[i]
            sel_b = b_comp[i] * flags[i] if sel_b is None else sel_b + b_comp[i] * flags[i]
            sel_c = c_comp[i] * flags[i] if sel_c is None else sel_c + c_comp[i] * flags[i]
            b.when(1 - visited).assert_eq(b_comp[i], c_comp[i])
            b.when(is_comp_eq).assert_zero(visited)
        b.assert_eq(cmp_b, sel_b)
        b.assert_eq(cmp_c, sel_c)
        b.when(1 - is_comp_eq).assert_eq(not_eq_inv * (cmp_b - cmp_c), is_real)
        _assert_bool(b, is_slt)
        _assert_bool(b, is_sltu)
        _assert_bool(b, is_real)
    return Chip("Lt", 0, 32, ev, local_only=True)


def _shift_right_chip() -> Chip:
    """ShiftRightChip::eval crates/core/machine/src/alu/sr/mod.rs:352-512 without its byte lookups (MSB of b,
    ShrCarry per byte, range checks) and its instruction receive; columns of ShiftRightCols."""
    def ev(b):
        bw = [b.main(2 + i) for i in range(4)]
        c0 = b.main(6)
        by_bits = [b.main(10 + i) for i in range(8)]
        by_bytes = [b.main(18 + i) for i in range(4)]
        byte_res = [b.main(22 + i) for i in range(8)]
        bit_res = [b.main(30 + i) for i in range(8)]
        carry = [b.main(38 + i) for i in range(8)]
        shifted = [b.main(46 + i) for i in range(8)]
        b_msb = b.main(54)
        c_bits = [b.main(55 + i) for i in range(8)]
        is_srl, is_ror, is_sra, is_real = b.main(63), b.main(64), b.main(65), b.main(66)
        total = c_bits[0]
        for i in range(1, 8):
            total = total + c_bits[i] * (1 << i)
        b.assert_eq(total, c0)
        nbits = c_bits[0] + c_bits[1] * 2 + c_bits[2] * 4
        for i in range(8):
            b.when(by_bits[i]).assert_eq(nbits, i)
        tot = by_bits[0]
        for x in by_bits[1:]:
            tot = tot + x
        b.assert_eq(tot, 1)
        nbytes = c_bits[3] + c_bits[4] * 2
        for i in range(4):
            b.when(by_bytes[i]).assert_eq(nbytes, i)
        tot = by_bytes[0]
        for x in by_bytes[1:]:
            tot = tot + x
        b.assert_eq(tot, 1)
        ext = list(bw) + [is_sra * b_msb * 0xFF + is_ror * bw[i] for i in range(4)]
        for k in range(4):
            for i in range(8 - k):
                b.when(by_bytes[k]).assert_eq(byte_res[i], ext[i + k])
        mult = by_bits[0] * (1 << 8)
        for i in range(1, 8):
            mult = mult + by_bits[i] * (1 << (8 - i))
        for i in range(7, -1, -1):
            v = shifted[i]
            if i + 1 < 8:
                v = v + carry[i + 1] * mult
            b.assert_eq(v, bit_res[i])
        for f in (is_srl, is_sra, is_ror, is_real, b_msb):
            _assert_bool(b, f)
        for f in by_bytes + by_bits + c_bits:
            _assert_bool(b, f)
        b.assert_eq(is_srl + is_sra + is_ror, is_real)
    return Chip("ShiftRight", 0, 67, ev, local_only=True)


def _bitwise_chip() -> Chip:
    """BitwiseChip::eval crates/core/machine/src/alu/bitwise/mod.rs:205-255: everything but the flag
    constraints is a byte lookup."""
    def ev(b):
        flags = [b.main(14 + i) for i in range(4)]
        for f in flags:
            _assert_bool(b, f)
        _assert_bool(b, flags[0] + flags[1] + flags[2] + flags[3])
    return Chip("Bitwise", 0, 18, ev, local_only=True)


def _clo_clz_chip() -> Chip:
    """CloClzChip::eval crates/core/machine/src/alu/clo_clz/mod.rs:187-281 without its lookups (a[0] < 33, the SRL
    send that pins the count) and its instruction receive; columns of CloClzCols."""
    def ev(b):
        a = [b.main(2 + i) for i in range(4)]
        bw = [b.main(6 + i) for i in range(4)]
        bb = [b.main(10 + i) for i in range(4)]
        is_bb_zero, is_clz, is_real = b.main(14), b.main(15), b.main(16)
        is_clo = is_real - is_clz
        for x, y in zip(bw, bb):
            b.when(is_clo).assert_eq(x + y, 255)
            b.when(is_clz).assert_eq(x, y)
        for i in (1, 2, 3):
            b.when(is_real).assert_zero(a[i])
        _assert_bool(b, is_bb_zero)
        b.when(is_bb_zero).assert_zero(bb[0] + bb[1] * (1 << 8) + bb[2] * (1 << 16) + bb[3] * (1 << 24))
        b.when(is_bb_zero).assert_zero(bb[3])
        b.when(is_bb_zero).assert_eq(a[0], 32)
        _assert_bool(b, is_clz)
        _assert_bool(b, is_real)
        b.when(is_clz).assert_eq(is_real, 1)
    return Chip("CloClz", 0, 17, ev, local_only=True)


def _word_value(w):
    """Word::reduce: the field element sum_i byte_i 256^i."""
    return w[0] + w[1] * (1 << 8) + w[2] * (1 << 16) + w[3] * (1 << 24)


def _range_check_word(b, value, rc, is_real):
    """KoalaBearWordRangeChecker::range_check crates/core/machine/src/operations/koala_bear_word.rs:51-110:
    rc = 8 bits of the most significant byte + the running conjunctions of its low 2..7 bits."""
    real = b.when(is_real)
    byte = rc[0]
    for i in range(8):
        real.assert_zero(rc[i] * (rc[i] - 1))
        if i:
            byte = byte + rc[i] * (1 << i)
    real.assert_eq(byte, value[3])
    real.assert_zero(rc[7])
    real.assert_eq(rc[8], rc[0] * rc[1])
    for j in range(5):
        real.assert_eq(rc[9 + j], rc[8 + j] * rc[2 + j])
    b.when(is_real).when(rc[13]).assert_zero(value[0] + value[1] + value[2])


def _branch_chip() -> Chip:
    """BranchChip::eval crates/core/machine/src/control_flow/branch/air.rs:20-211 without its lookups (instruction
    receive, ADD for the target, SLT for a_lt_b / a_gt_b, byte range checks); columns of BranchColumns."""
    def ev(b):
        next_pc = [b.main(1 + i) for i in range(4)]
        next_rc = [b.main(5 + i) for i in range(14)]
        target = [b.main(19 + i) for i in range(4)]
        nn_pc = [b.main(23 + i) for i in range(4)]
        nn_rc = [b.main(27 + i) for i in range(14)]
        is_beq, is_bne, is_bltz, is_blez, is_bgtz, is_bgez = (b.main(53 + i) for i in range(6))
        is_branching, a_gt_b, a_lt_b = b.main(59), b.main(60), b.main(61)
        flags = [is_beq, is_bne, is_bltz, is_bgez, is_blez, is_bgtz]
        for f in flags:
            _assert_bool(b, f)
        is_real = is_beq + is_bne + is_bltz + is_bgez + is_blez + is_bgtz
        _assert_bool(b, is_real)
        _range_check_word(b, next_pc, next_rc, is_real)
        _range_check_word(b, nn_pc, nn_rc, is_real)
        b.when(is_real).when(1 - is_branching).assert_eq(_word_value(next_pc) + 4, _word_value(nn_pc))
        for i in range(4):
            b.when(is_real).when(is_branching).assert_eq(target[i], nn_pc[i])
        b.when(1 - is_real).assert_zero(is_branching)
        b.when(is_real).assert_zero(is_branching * (is_branching - 1))
        either = a_gt_b + a_lt_b
        b.when(is_beq * is_branching).assert_zero(either)
        b.when(is_beq).when(1 - is_branching).assert_eq(either, 1)
        b.when(is_bne * is_branching).assert_eq(either, 1)
        b.when(is_bne).when(1 - is_branching).assert_zero(either)
        b.when(is_bltz * is_branching).assert_eq(a_lt_b, 1)
        b.when(is_bltz).when(1 - is_branching).assert_zero(a_lt_b)
        b.when(is_blez * is_branching).assert_zero(a_gt_b)
        b.when(is_blez).when(1 - is_branching).assert_eq(a_gt_b, 1)
        b.when(is_bgtz * is_branching).assert_eq(a_gt_b, 1)
        b.when(is_bgtz).when(1 - is_branching).assert_zero(a_gt_b)
        b.when(is_bgez * is_branching).assert_zero(a_lt_b)
        b.when(is_bgez).when(1 - is_branching).assert_eq(a_lt_b, 1)
    return Chip("Branch", 0, 62, ev, local_only=True)


def _jump_chip() -> Chip:
    """JumpChip::eval crates/core/machine/src/control_flow/jump/air.rs:20-115 without its lookups (instruction
    receive, ADD for the pc-relative target); columns of JumpColumns."""
    def ev(b):
        next_pc = [b.main(1 + i) for i in range(4)]
        next_rc = [b.main(5 + i) for i in range(14)]
        nn_pc = [b.main(19 + i) for i in range(4)]
        nn_rc = [b.main(23 + i) for i in range(14)]
        op_a = [b.main(37 + i) for i in range(4)]
        op_b = [b.main(41 + i) for i in range(4)]
        is_jump, is_jumpi, is_jumpdirect = b.main(49), b.main(50), b.main(51)
        a_rc = [b.main(52 + i) for i in range(14)]
        for f in (is_jump, is_jumpi, is_jumpdirect):
            _assert_bool(b, f)
        is_real = is_jump + is_jumpi + is_jumpdirect
        _assert_bool(b, is_real)
        b.when(is_real).assert_eq(_word_value(op_a), _word_value(next_pc) + 4)
        _range_check_word(b, op_a, a_rc, is_real)
        _range_check_word(b, next_pc, next_rc, is_real)
        _range_check_word(b, nn_pc, nn_rc, is_real)
        for i in range(4):
            b.when(is_jump + is_jumpi).assert_eq(nn_pc[i], op_b[i])
    return Chip("Jump", 0, 66, ev, local_only=True)


def _mov_cond_chip() -> Chip:
    """MovCondChip::eval crates/core/machine/src/misc/mov_cond/mod.rs:170-255 with IsZeroWordOperation::eval
    (operations/is_zero_word.rs:49-82, is_zero.rs:42-66), without the instruction receive; columns of MovCondCols."""
    def ev(b):
        op_a = [b.main(2 + i) for i in range(4)]
        prev_a = [b.main(6 + i) for i in range(4)]
        op_b = [b.main(10 + i) for i in range(4)]
        op_c = [b.main(14 + i) for i in range(4)]
        inv = [b.main(18 + 2 * i) for i in range(4)]
        res = [b.main(19 + 2 * i) for i in range(4)]
        lower, upper, result = b.main(26), b.main(27), b.main(28)
        is_mne, is_meq, is_wsbh = b.main(29), b.main(30), b.main(31)
        is_real = is_mne + is_meq + is_wsbh
        real = b.when(is_real)
        for i in range(4):
            real.assert_eq(1 - inv[i] * op_c[i], res[i])
            real.assert_zero(res[i] * (res[i] - 1))
            b.when(is_real).when(res[i]).assert_zero(op_c[i])
        _assert_bool(b, is_real)
        for f in (lower, upper, result):
            real.assert_zero(f * (f - 1))
        real.assert_eq(lower, res[0] * res[1])
        real.assert_eq(upper, res[2] * res[3])
        real.assert_eq(result, lower * upper)
        for i in range(4):
            b.when(is_meq).when(result).assert_eq(op_a[i], op_b[i])
            b.when(is_meq).when(1 - result).assert_eq(op_a[i], prev_a[i])
            b.when(is_mne).when(1 - result).assert_eq(op_a[i], op_b[i])
            b.when(is_mne).when(result).assert_eq(op_a[i], prev_a[i])
        for i, j in ((0, 1), (1, 0), (2, 3), (3, 2)):
            b.when(is_wsbh).assert_eq(op_a[i], op_b[j])
        for i in range(4):
            b.when(is_wsbh).assert_zero(prev_a[i])
        for f in (is_mne, is_meq, is_wsbh):
            _assert_bool(b, f)
    return Chip("MovCond", 0, 32, ev, local_only=True)


def alu_case(traces: dict, *, with_lookup_pair: bool = True, **kw) -> ShardCase:
    """`traces`: {"AddSub": rows, "ShiftLeft": rows[, "Lt" / "ShiftRight" / "Bitwise" / "CloClz": rows]} in canonical
    form, as produced by trace generation.
    with_lookup_pair adds the Fibonacci/Sink pair so that the shard also has permutation traces (the two
    ALU chips alone have no lookups here)."""
    optional = {"Lt": _lt_chip, "ShiftRight": _shift_right_chip, "Bitwise": _bitwise_chip, "CloClz": _clo_clz_chip,
                "Branch": _branch_chip, "Jump": _jump_chip, "MovCond": _mov_cond_chip}
    chips = [_add_sub_chip(), _shift_left_chip()] + [make() for name, make in optional.items() if name in traces]
    traces = dict(traces)
    pv = np.zeros(8, dtype=np.uint32)
    if with_lookup_pair:
        n = 1 << 5
        a, b_ = 0, 1
        rows = np.empty((n, 2), dtype=np.uint32)
        for i in range(n):
            rows[i] = (a, b_)
            a, b_ = b_, (a + b_) % P
        pv[1], pv[2], pv[3] = 0, 1, rows[-1, 1]
        sink = np.zeros((n, 3), dtype=np.uint32)
        sink[:, :2] = rows[::-1]
        sink[:, 2] = 1
        chips += [_fib_chip(), _sink_chip()]
        traces["Fibonacci"], traces["Sink"] = rows, sink
    machine = Machine(chips, num_pv_elts=4, num_queries=kw.get("num_queries", 8), pow_bits=kw.get("pow_bits", 4),
                      log_blowup=kw.get("log_blowup", 1))
    cycles = sum(int(traces[k].shape[0]) for k in ("AddSub", "ShiftLeft", "Lt", "ShiftRight", "Bitwise", "CloClz", "Branch", "Jump", "MovCond") if k in traces)
    return ShardCase(machine, {}, traces, pv, cycles)


# ---- the Global chip (crates/core/machine/src/global/mod.rs) -------------------------------------------------------------------
# CURVE_CUMULATIVE_SUM_START_{X,Y}, crates/stark/src/septic_digest.rs:9-14
GLOBAL_START_POINT = (637514027, 1595065213, 1998064738, 72333738, 1211544370, 822986770, 1518535784,
                      1604177449, 90440090, 259343427, 140470264, 1162099742, 941559812, 1064053343)


def _sep_mul(a, b):
    """Product in F_p[z] / (z^7 + 2z - 8) of two coefficient lists of expressions (SepticExtension::mul,
    crates/stark/src/septic_extension.rs:306-323)."""
    r = [None] * 13
    for i in range(7):
        for j in range(7):
            t = a[i] * b[j]
            r[i + j] = t if r[i + j] is None else r[i + j] + t
    out = r[:7]
    for i in range(7, 13):
        out[i - 7] = out[i - 7] + r[i] * 8
        out[i - 6] = out[i - 6] - r[i] * 2
    return out


def _sep_add(a, b):
    return [x + y for x, y in zip(a, b)]


def _sep_sub(a, b):
    return [x - y for x, y in zip(a, b)]


def _global_chip(with_receive: bool = False) -> Chip:
    """GlobalChip::eval crates/core/machine/src/global/mod.rs:216-276 with GlobalLookupOperation::eval_single_digest
    (operations/global_lookup.rs:93-175) and GlobalAccumulationOperation::eval_accumulation
    (operations/global_accumulation.rs:129-224), N = 1; columns of GlobalCols (global/mod.rs:53-63).  with_receive: the receive
    of (message, is_send, is_receive, kind) from the tables that emit global lookups (global/mod.rs:232-250); left out: the
    U16Range byte lookup of message[0].
    The chip commits in the global scope: its last fourteen columns are the shard's cumulative sum."""
    def ev(b):
        m = [b.main(i) for i in range(7)]
        kind = b.main(7)
        offset_bits = [b.main(8 + i) for i in range(8)]
        x = [b.main(16 + i) for i in range(7)]
        y = [b.main(23 + i) for i in range(7)]
        y6_bits = [b.main(30 + i) for i in range(30)]
        witness, is_receive, is_send, is_real = b.main(60), b.main(61), b.main(62), b.main(63)
        init_x, init_y = [b.main(64 + i) for i in range(7)], [b.main(71 + i) for i in range(7)]
        checker = [b.main(78 + i) for i in range(7)]
        cum_x, cum_y = [b.main(85 + i) for i in range(7)], [b.main(92 + i) for i in range(7)]
        next_real = b.main(63, next=True)
        next_init_x, next_init_y = [b.main(64 + i, next=True) for i in range(7)], [b.main(71 + i, next=True) for i in range(7)]
        real = b.when(is_real)
        if with_receive:
            b.receive(KIND_GLOBAL, m + [is_send, is_receive, kind], is_real)
        # eval_single_digest
        _assert_bool(b, is_real)
        offset = offset_bits[0]
        for i in range(8):
            _assert_bool(b, offset_bits[i])
            if i:
                offset = offset + offset_bits[i] * (1 << i)
        real.assert_eq(x[0], m[0] + kind * 65536)
        for i in range(1, 6):
            real.assert_eq(x[i], m[i])
        real.assert_eq(x[6], m[6] * 256 + offset)
        y2 = _sep_mul(y, y)
        x3 = _sep_mul(_sep_mul(x, x), x)
        rhs = list(x3)
        # + 3z * x - 3: z * (x0 .. x6) = (8 x6, x0 - 2 x6, x1, .. x5)
        zx = [x[6] * 8, x[0] - x[6] * 2] + [x[i] for i in range(1, 6)]
        rhs = [rhs[i] + zx[i] * 3 for i in range(7)]
        rhs[0] = rhs[0] - 3
        for i in range(7):
            b.assert_eq(y2[i], rhs[i])
        y6_value, top = y6_bits[0], None
        for i in range(30):
            _assert_bool(b, y6_bits[i])
            if i:
                y6_value = y6_value + y6_bits[i] * (1 << i)
            if i >= 23:
                top = y6_bits[i] if top is None else top + y6_bits[i]
        real.assert_eq(witness * (top - 7), 1)
        b.when(is_receive).assert_eq(y[6], y6_value + 1)
        b.when(is_send).assert_eq(y[6], y6_value + ((1 << 30) - (1 << 23) + 1))
        # eval_accumulation
        b.when_transition().when(1 - is_real).assert_zero(next_real)
        first = b.when_first_row()
        for i in range(7):
            first.assert_eq(init_x[i], GLOBAL_START_POINT[i])
            first.assert_eq(init_y[i], GLOBAL_START_POINT[7 + i])
        dx, dy = _sep_sub(x, init_x), _sep_sub(y, init_y)
        chk_x = _sep_sub(_sep_mul(_sep_add(_sep_add(init_x, x), cum_x), _sep_mul(dx, dx)), _sep_mul(dy, dy))
        chk_y = _sep_sub(_sep_mul(_sep_add(init_y, cum_y), dx), _sep_mul(dy, _sep_sub(init_x, cum_x)))
        not_real = b.when(1 - is_real)
        trans = b.when_transition()
        for i in range(7):
            b.assert_eq(chk_x[i], checker[i])
            real.assert_zero(checker[i])
            real.assert_zero(chk_y[i])
            not_real.assert_eq(init_x[i], cum_x[i])
            not_real.assert_eq(init_y[i], cum_y[i])
            trans.assert_eq(cum_x[i], next_init_x[i])
            trans.assert_eq(cum_y[i], next_init_y[i])
    return Chip("Global", 0, 99, ev, global_scope=True)


def global_case(global_rows: np.ndarray, **kw) -> ShardCase:
    """A shard with the Global table (canonical rows as trace generation produces them) under the chip's restated constraints,
    next to the Fibonacci / Sink pair so that the shard also has permutation traces."""
    n = 1 << 5
    a, b_ = 0, 1
    rows = np.empty((n, 2), dtype=np.uint32)
    for i in range(n):
        rows[i] = (a, b_)
        a, b_ = b_, (a + b_) % P
    pv = np.zeros(8, dtype=np.uint32)
    pv[1], pv[2], pv[3] = 0, 1, rows[-1, 1]
    sink = np.zeros((n, 3), dtype=np.uint32)
    sink[:, :2] = rows[::-1]
    sink[:, 2] = 1
    machine = Machine([_global_chip(), _fib_chip(), _sink_chip()], num_pv_elts=4, num_queries=kw.get("num_queries", 8),
                      pow_bits=kw.get("pow_bits", 4), log_blowup=kw.get("log_blowup", 1))
    return ShardCase(machine, {}, {"Global": global_rows, "Fibonacci": rows, "Sink": sink}, pv, int(global_rows.shape[0]))


def _memory_global_chip(finalize: bool, pv_prev: int, pv_last: int) -> Chip:
    """MemoryGlobalChip::eval crates/core/machine/src/memory/global.rs:248-445 with KoalaBearBitDecomposition::range_check
    (operations/koala_bear_range.rs:50-117), AssertLtColsBits::eval (operations/cmp.rs:322-391) and IsZeroOperation::eval
    (operations/is_zero.rs:42-71); columns of MemoryInitCols (global.rs:210-245).  pv_prev / pv_last: where the 32 bits of
    previous_{init,finalize}_addr_bits and last_{init,finalize}_addr_bits start in the public values."""
    def ev(b):
        shard, timestamp, addr = b.main(0), b.main(1), b.main(2)
        lt = [b.main(3 + i) for i in range(32)]
        bits = [b.main(35 + i) for i in range(32)]
        ands = [b.main(67 + i) for i in range(6)]
        value = [b.main(73 + i) for i in range(32)]
        is_real, is_next_comp, inverse, prev_zero, is_first_comp, is_last_addr = (b.main(105 + i) for i in range(6))
        n_lt = [b.main(3 + i, next=True) for i in range(32)]
        n_bits = [b.main(35 + i, next=True) for i in range(32)]
        n_real, n_next_comp = b.main(105, next=True), b.main(106, next=True)
        prev_bits = [b.pub(pv_prev + i) for i in range(32)]
        last_bits = [b.pub(pv_last + i) for i in range(32)]
        _assert_bool(b, is_real)
        for v in value:
            _assert_bool(b, v)
        byte = [sum((value[8 * k + i] * (1 << i) for i in range(1, 8)), value[8 * k]) for k in range(4)]
        if finalize:
            b.send(KIND_GLOBAL, [shard, timestamp, addr] + byte + [is_real * 0, is_real * 1, KIND_MEMORY], is_real)
        else:
            b.send(KIND_GLOBAL, [0, 0, addr] + byte + [is_real * 1, is_real * 0, KIND_MEMORY], is_real)
        # KoalaBearBitDecomposition::range_check(addr, addr_bits, is_real)
        real = b.when(is_real)
        for x in bits:
            real.assert_zero(x * (x - 1))
        real.assert_eq(sum((bits[i] * ((1 << i) % P) for i in range(1, 32)), bits[0]), addr)
        top = bits[24:32]
        real.assert_zero(top[7])
        real.assert_eq(ands[0], top[0] * top[1])
        for i in range(1, 6):
            real.assert_eq(ands[i], ands[i - 1] * top[i + 1])
        real.when(ands[5]).assert_zero(sum(bits[1:24], bits[0]))

        def assert_lt(flags, a_bits, b_bits, cond):             # AssertLtColsBits::eval
            for f in flags:
                _assert_bool(b, f)
            on = b.when(cond)
            on.assert_eq(sum(flags[1:], flags[0]), 1)
            visited, a_cmp, b_cmp = None, None, None
            for i in reversed(range(32)):
                visited = flags[i] if visited is None else visited + flags[i]
                a_cmp = a_bits[i] * flags[i] if a_cmp is None else a_cmp + a_bits[i] * flags[i]
                b_cmp = b_bits[i] * flags[i] if b_cmp is None else b_cmp + b_bits[i] * flags[i]
                on.when(1 - visited).assert_eq(a_bits[i], b_bits[i])
            on.assert_zero(a_cmp)
            on.assert_eq(b_cmp, 1)

        b.when_transition().assert_eq(n_next_comp, n_real)
        assert_lt(n_lt, bits, n_bits, n_next_comp)
        b.when_transition().when(1 - is_real).assert_zero(n_real)
        prev_addr = sum((prev_bits[i] * ((1 << i) % P) for i in range(1, 32)), prev_bits[0])
        first = b.when_first_row()
        first.assert_eq(1 - inverse * prev_addr, prev_zero)      # IsZeroOperation::eval(prev_addr, is_prev_addr_zero, is_first_row)
        first.assert_zero(prev_zero * (prev_zero - 1))
        first.when(prev_zero).assert_zero(prev_addr)
        _assert_bool(b, is_first_comp)
        first.assert_eq(is_first_comp, 1 - prev_zero)
        first.assert_eq(is_real, 1)
        assert_lt(lt, prev_bits, bits, is_first_comp)
        first.when(prev_zero).assert_zero(addr)
        first.when(prev_zero).assert_eq(n_real, 1)
        first.when(prev_zero).assert_eq(n_next_comp, 1)
        if not finalize:
            real.assert_eq(timestamp, 1)
        for v in value:
            first.when(1 - is_first_comp).assert_zero(v)
        b.when_transition().assert_eq(is_last_addr, is_real * (1 - n_real))
        for x, pub in zip(bits, last_bits):
            b.when_last_row().when(is_real).assert_eq(x, pub)
            b.when_transition().when(is_last_addr).assert_eq(x, pub)
    return Chip("MemoryGlobalFinalize" if finalize else "MemoryGlobalInit", 0, 111, ev)


def _syscall_chip(precompile: bool) -> Chip:
    """SyscallChip::eval crates/core/machine/src/syscall/chip.rs:304-497; columns of SyscallCols (chip.rs:71-107).  The two
    cross-shard lookups every real row sends to the Global table - (shard, clk, syscall_id, arg half-words) of kind Syscall and
    (shard, clk, syscall_id, result half-words, 0, 0) of kind SyscallResult, as a send in a core shard and as a receive in a
    precompile shard - are stated; left out: the four U16Range byte lookups and the syscall / syscall-result lookups with the
    SyscallInstrs table or the precompile tables (their other ends are tables this machine does not have)."""
    def ev(b):
        shard, clk, syscall_id, a1_lo, a1_hi, a2_lo, a2_hi, r_lo, r_hi, is_linux, is_real = (b.main(i) for i in range(11))
        _assert_bool(b, is_real)
        _assert_bool(b, is_linux)
        b.when(1 - is_real).assert_zero(is_linux)
        b.when(1 - is_linux).assert_zero(r_lo)
        b.when(1 - is_linux).assert_zero(r_hi)
        is_send, is_receive = (is_real * 0, is_real * 1) if precompile else (is_real * 1, is_real * 0)
        b.send(KIND_GLOBAL, [shard, clk, syscall_id, a1_lo, a1_hi, a2_lo, a2_hi, is_send, is_receive, KIND_SYSCALL], is_real)
        b.send(KIND_GLOBAL, [shard, clk, syscall_id, r_lo, r_hi, 0, 0, is_send, is_receive, KIND_SYSCALL_RESULT], is_real)
    return Chip("SyscallPrecompile" if precompile else "SyscallCore", 0, 11, ev)


def memory_global_case(init_rows: np.ndarray, finalize_rows: np.ndarray, global_rows: np.ndarray, previous_init_addr: int,
                       previous_finalize_addr: int, syscall_rows: np.ndarray | None = None, **kw) -> ShardCase:
    """A shard of the three tables that carry memory across shards, under their restated constraints AND the lookup that ties
    them: MemoryGlobalInit / MemoryGlobalFinalize send (shard, timestamp, addr, value bytes, is_send, is_receive, Memory) for
    every real row, Global receives its messages.  Public values: the bits of previous_init_addr, last_init_addr,
    previous_finalize_addr, last_finalize_addr (the last addresses read off the tables)."""
    def last_addr(rows):
        real = rows[rows[:, 105] == 1]
        return int(real[-1, 2]) if len(real) else 0
    pv = np.zeros(128, dtype=np.uint32)
    for k, a in enumerate((previous_init_addr, last_addr(init_rows), previous_finalize_addr, last_addr(finalize_rows))):
        pv[32 * k: 32 * k + 32] = [(a >> i) & 1 for i in range(32)]
    chips = [_memory_global_chip(False, 0, 32), _memory_global_chip(True, 64, 96), _global_chip(with_receive=True)]
    if syscall_rows is not None:            # a core shard's syscall table: two more global lookups per row
        chips.append(_syscall_chip(False))
    machine = Machine(chips, num_pv_elts=128, num_queries=kw.get("num_queries", 8), pow_bits=kw.get("pow_bits", 4),
                      log_blowup=kw.get("log_blowup", 1))
    traces = {"MemoryGlobalInit": init_rows, "MemoryGlobalFinalize": finalize_rows, "Global": global_rows}
    if syscall_rows is not None:
        traces["SyscallCore"] = syscall_rows
    return ShardCase(machine, {}, traces, pv, int(global_rows.shape[0]))


def _eval_is_zero_word(b, a, cols, is_real):
    """IsZeroWordOperation::eval (crates/core/machine/src/operations/is_zero_word.rs:49-82) over four expressions `a`; cols = the
    operation's eleven columns {is_zero_byte[4]{inverse, result}, is_lower_half_zero, is_upper_half_zero, result}."""
    on = b.when(is_real)
    for i in range(4):
        inverse, result = cols[2 * i], cols[2 * i + 1]
        on.assert_eq(1 - inverse * a[i], result)               # IsZeroOperation::eval, is_zero.rs:42-71
        on.assert_zero(result * (result - 1))
        on.when(result).assert_zero(a[i])
    _assert_bool(b, is_real)
    lower, upper, result = cols[8], cols[9], cols[10]
    for x in (lower, upper, result):
        on.assert_zero(x * (x - 1))
    on.assert_eq(lower, cols[1] * cols[3])
    on.assert_eq(upper, cols[5] * cols[7])
    on.assert_eq(result, lower * upper)


def _div_rem_chip() -> Chip:
    """DivRemChip::eval crates/core/machine/src/alu/divrem/mod.rs:390-795; columns of DivRemCols (mod.rs:109-204).  Left out (lookups
    whose other ends are other tables): the MULT / MULTU send for c * quotient, the ADD sends for the absolute values, the SLTU
    send for |remainder| < max(|c|, 1), the MSB and range byte lookups, the instruction receives and the HI register's memory
    access."""
    def ev(b):
        word = lambda c0: [b.main(c0 + i) for i in range(4)]
        bb, c, quotient, remainder, abs_remainder, abs_c, max_abs_c_or_1 = (word(2 + 4 * k) for k in range(7))
        ctq = [b.main(30 + i) for i in range(8)]
        carry = [b.main(38 + i) for i in range(8)]
        is_c_0 = [b.main(46 + i) for i in range(11)]
        is_div, is_divu, is_mod, is_modu, is_overflow = (b.main(57 + i) for i in range(5))
        ovf_b = [b.main(62 + i) for i in range(11)]
        ovf_c = [b.main(73 + i) for i in range(11)]
        b_msb, rem_msb, c_msb, b_neg, rem_neg, c_neg, multiplicity = (b.main(84 + i) for i in range(7))
        hi_value = word(95)                                     # op_hi_access.value(): prev_value[4], then the access columns' value
        is_real = is_div + is_divu + is_mod + is_modu
        signed = is_div + is_mod
        for msb, neg in ((b_msb, b_neg), (rem_msb, rem_neg), (c_msb, c_neg)):
            b.assert_eq(msb * signed, neg)
        int_min, minus_one = [0, 0, 0, 0x80], [0xFF] * 4
        _eval_is_zero_word(b, [bb[i] - int_min[i] for i in range(4)], ovf_b, is_real)
        _eval_is_zero_word(b, [c[i] - minus_one[i] for i in range(4)], ovf_c, is_real)
        b.assert_eq(is_overflow, ovf_b[10] * ovf_c[10] * signed)
        # c * quotient + remainder = b, byte by byte with carries, the remainder sign-extended
        total = []
        for i in range(8):
            t = ctq[i] + (remainder[i] if i < 4 else rem_neg * 255) - carry[i] * 256
            total.append(t + carry[i - 1] if i else t)
        for i in range(4):
            b.assert_eq(bb[i], total[i])
        for i in range(4, 8):
            b.when(1 - is_overflow).when(b_neg).assert_eq(total[i], 255)
            b.when(1 - is_overflow).when(1 - b_neg).assert_zero(total[i])
            b.when(is_overflow).assert_zero(total[i])
        rem_sum = remainder[0] + remainder[1] + remainder[2] + remainder[3]
        b.when(rem_neg).assert_eq(b_neg, 1)
        b.when(rem_sum).when(1 - rem_neg).assert_zero(b_neg)
        _eval_is_zero_word(b, c, is_c_0, is_real)
        for i in range(4):
            b.when(is_c_0[10]).assert_eq(quotient[i], 255)
            b.when(1 - c_neg).assert_eq(c[i], abs_c[i])
            b.when(1 - rem_neg).assert_eq(remainder[i], abs_remainder[i])
        b.when(is_real).assert_eq(max_abs_c_or_1[0], is_c_0[10] + (1 - is_c_0[10]) * abs_c[0])
        for i in range(1, 4):
            b.when(is_real).assert_eq(max_abs_c_or_1[i], (1 - is_c_0[10]) * abs_c[i])
        b.assert_eq((1 - is_c_0[10]) * is_real, multiplicity)
        for x in carry:
            _assert_bool(b, x)
        for x in (is_div, is_divu, is_mod, is_modu, is_overflow, b_msb, rem_msb, c_msb, b_neg, rem_neg, c_neg):
            _assert_bool(b, x)
        b.when(is_real).assert_eq(is_divu + is_div + is_mod + is_modu, 1)
        for i in range(4):
            b.when(is_div + is_divu).assert_eq(remainder[i], hi_value[i])
    return Chip("DivRem", 0, 106, ev, local_only=True)


def chips_case(traces: dict, **kw) -> ShardCase:
    """Tables of the chips whose Air::eval is restated here one by one ({"DivRem": rows, ...}, canonical rows from trace
    generation) next to the Fibonacci / Sink pair, so that the shard also has permutation traces."""
    makers = {"DivRem": _div_rem_chip}
    n = 1 << 5
    a, b_ = 0, 1
    rows = np.empty((n, 2), dtype=np.uint32)
    for i in range(n):
        rows[i] = (a, b_)
        a, b_ = b_, (a + b_) % P
    pv = np.zeros(8, dtype=np.uint32)
    pv[1], pv[2], pv[3] = 0, 1, rows[-1, 1]
    sink = np.zeros((n, 3), dtype=np.uint32)
    sink[:, :2] = rows[::-1]
    sink[:, 2] = 1
    machine = Machine([makers[k]() for k in traces] + [_fib_chip(), _sink_chip()], num_pv_elts=4, num_queries=kw.get("num_queries", 8),
                      pow_bits=kw.get("pow_bits", 4), log_blowup=kw.get("log_blowup", 1))
    return ShardCase(machine, {}, {**traces, "Fibonacci": rows, "Sink": sink}, pv, sum(int(v.shape[0]) for v in traces.values()))


def syscall_precompile_case(syscall_rows: np.ndarray, global_rows: np.ndarray, **kw) -> ShardCase:
    """A precompile shard's cross-shard tables: SyscallPrecompile (SyscallChip::eval with shard_kind Precompile: its two global
    lookups per row are RECEIVES) and Global, tied by the real lookup."""
    machine = Machine([_syscall_chip(True), _global_chip(with_receive=True)], num_pv_elts=4, num_queries=kw.get("num_queries", 8),
                      pow_bits=kw.get("pow_bits", 4), log_blowup=kw.get("log_blowup", 1))
    return ShardCase(machine, {}, {"SyscallPrecompile": syscall_rows, "Global": global_rows}, np.zeros(8, dtype=np.uint32),
                     int(global_rows.shape[0]))


def _syscall_instrs_chip() -> Chip:
    """SyscallInstrsChip::eval crates/core/machine/src/syscall/instructions/air.rs:28-446; columns of SyscallInstrColumns
    (columns.rs:11-59).  Public values as this machine lays them out: [0, 32) committed_value_digest (eight words as bytes),
    [32, 40) deferred_proofs_digest, 40 exit_code.  Left out (lookups whose other ends are other tables): the instruction receive
    from the Cpu table, the syscall / syscall-result sends to SyscallCore and SysLinux."""
    def ev(b):
        pc, next_pc, shard, clk, num_extra_cycles, is_halt, is_sys_linux = (b.main(i) for i in range(7))
        a1_zero = (b.main(7), b.main(8))
        syscall_id_col = b.main(9)
        word = lambda c0: [b.main(c0 + i) for i in range(4)]
        op_a, op_b, op_c, prev_a = word(10), word(14), word(18), word(22)
        iz = lambda c0: (b.main(c0), b.main(c0 + 1))
        is_enter, is_hint, halt_chk, exit_chk, is_commit, is_deferred = (iz(26 + 2 * k) for k in range(6))
        bitmap = [b.main(38 + i) for i in range(8)]
        b_rc, c_rc = [b.main(46 + i) for i in range(14)], [b.main(60 + i) for i in range(14)]
        b_check, c_check, is_real = b.main(74), b.main(75), b.main(76)
        digest = [[b.pub(4 * w + k) for k in range(4)] for w in range(8)]
        deferred = [b.pub(32 + i) for i in range(8)]
        exit_code = b.pub(40)
        real = b.when(is_real)
        reduce = lambda w: w[0] + w[1] * 256 + w[2] * 65536 + w[3] * (1 << 24)

        def is_zero(a, cols):                                   # IsZeroOperation::eval(a, cols, is_real), is_zero.rs:42-71
            inverse, result = cols
            real.assert_eq(1 - inverse * a, result)
            real.assert_zero(result * (result - 1))
            real.when(result).assert_zero(a)
            return result

        _assert_bool(b, is_real)
        sid = prev_a[0] + prev_a[1] * 256
        # eval_is_halt_syscall
        halt, exit_group = is_zero(sid - 0x00, halt_chk), is_zero(sid - 4246, exit_chk)
        b.assert_eq(is_halt, (halt + exit_group) * is_real)
        b.assert_eq(num_extra_cycles, prev_a[3] * is_real)
        # eval_syscall
        send_to_table = is_sys_linux + prev_a[2]
        _assert_bool(b, prev_a[2])
        _assert_bool(b, is_sys_linux)
        _assert_bool(b, send_to_table)
        real.assert_eq(is_sys_linux, 1 - is_zero(prev_a[1], a1_zero))
        b.when(1 - is_real).assert_zero(send_to_table)
        _assert_bool(b, b_check)
        _assert_bool(b, c_check)
        b.when(send_to_table).assert_eq(b_check, 1)
        b.when(is_halt).assert_eq(b_check, 1)
        b.when(send_to_table).assert_eq(c_check, 1)
        b.when(is_deferred[1]).assert_eq(c_check, 1)
        b.when(1 - is_real).assert_zero(b_check)
        b.when(1 - is_real).assert_zero(c_check)
        _range_check_word(b, op_b, b_rc, b_check)
        _range_check_word(b, op_c, c_rc, c_check)
        enter = is_zero(sid - 0x03, is_enter)
        real.when(1 - enter).assert_eq(syscall_id_col, sid)
        real.when(enter).assert_eq(syscall_id_col, 0x04)          # EXIT_UNCONSTRAINED
        hint = is_zero(sid - 0xF0, is_hint)
        for i in range(4):
            real.when(enter).assert_zero(op_a[i])
            real.when(1 - (enter + hint + is_sys_linux)).assert_eq(op_a[i], prev_a[i])
        # eval_commit
        commit, commit_deferred = is_zero(sid - 0x10, is_commit), is_zero(sid - 0x1A, is_deferred)
        for bit in bitmap:
            real.assert_zero(bit * (bit - 1))
        bitmap_sum = sum(bitmap[1:], bitmap[0])
        real.when(commit + commit_deferred).assert_eq(bitmap_sum, 1)
        real.when(1 - (commit + commit_deferred)).assert_zero(bitmap_sum)
        for i, bit in enumerate(bitmap):
            real.when(bit).assert_eq(op_b[0], i)
        for i in range(1, 4):
            real.when(commit + commit_deferred).assert_zero(op_b[i])
        for k in range(4):
            real.when(commit).assert_eq(sum((bitmap[w] * digest[w][k] for w in range(1, 8)), bitmap[0] * digest[0][k]), op_c[k])
        real.when(commit_deferred).assert_eq(sum((bitmap[w] * deferred[w] for w in range(1, 8)), bitmap[0] * deferred[0]), reduce(op_c))
        # eval_halt_unimpl
        b.when(is_halt).assert_zero(next_pc)
        b.when(is_halt).assert_eq(reduce(op_b), exit_code)
    return Chip("SyscallInstrs", 0, 77, ev, local_only=True)


def syscall_instrs_case(rows: np.ndarray, digest: np.ndarray, deferred: np.ndarray, exit_code: int, **kw) -> ShardCase:
    """The SyscallInstrs table under its restated constraints, with the public values its COMMIT / COMMIT_DEFERRED_PROOFS / HALT
    rows are checked against."""
    pv = np.zeros(48, dtype=np.uint32)
    for w in range(8):
        pv[4 * w: 4 * w + 4] = [(int(digest[w]) >> (8 * k)) & 0xFF for k in range(4)]
    pv[32:40] = deferred
    pv[40] = exit_code
    machine = Machine([_syscall_instrs_chip()], num_pv_elts=41, num_queries=kw.get("num_queries", 8), pow_bits=kw.get("pow_bits", 4),
                      log_blowup=kw.get("log_blowup", 1))
    return ShardCase(machine, {}, {"SyscallInstrs": rows}, pv, int(rows.shape[0]))

"""Chips as data: a small symbolic AIR builder that records what the reference's Rust chips
express through `AirBuilder` calls, and serialises a machine into the "ZKMD" descriptor the
C ABI (`zkb200_machine_create`, include/zkb200.h) consumes.

Mirrors, on the host side:
  * `AirBuilder::{main, is_first_row, is_last_row, is_transition, when*, assert_zero, assert_eq}`
    and `PairBuilder::preprocessed`, `AirBuilderWithPublicValues::public_values`
    (crates/stark/src/folder.rs:52-149);
  * `LookupBuilder` send/receive recording with `VirtualPairCol` linear combinations
    (crates/stark/src/lookup/lookup.rs:10-19, crates/stark/src/chip.rs:66-88);
  * `Chip::new`'s degree rule `log_quotient_degree = log2_ceil(max(deg, 3 if lookups) - 1)`
    (crates/stark/src/chip.rs:72-87) with p3 `SymbolicExpression::degree_multiple`
    (variables and is_first/is_last have degree 1, is_transition and constants 0).

In a Rust deployment this module is replaced by a `SymbolicAirBuilder` walk
(crates/stark/src/machine.rs:377-389) that emits the same descriptor; see INTEGRATION.md.
"""
from __future__ import annotations

import numpy as np

P = 0x7F000001

# node opcodes of the ZKMD expression DAG
N_CONST, N_MAIN, N_PREP, N_PUB, N_IS_FIRST, N_IS_LAST, N_IS_TRANS, N_ADD, N_SUB, N_MUL, N_NEG = range(11)

# LookupKind (crates/stark/src/lookup/lookup.rs:23-49)
KIND_MEMORY, KIND_PROGRAM, KIND_INSTRUCTION, KIND_BYTE, KIND_RANGE, KIND_SYSCALL, KIND_GLOBAL, KIND_SYSCALL_RESULT = range(1, 9)

SCOPE_LOCAL, SCOPE_GLOBAL = 0, 1


class Expr:
    __slots__ = ("b", "id", "deg", "lin")

    def __init__(self, b, id_, deg, lin):
        self.b, self.id, self.deg, self.lin = b, id_, deg, lin

    def _lift(self, o):
        return o if isinstance(o, Expr) else self.b.const(o)

    def __add__(self, o):
        o = self._lift(o)
        return self.b._bin(N_ADD, self, o)

    __radd__ = __add__

    def __sub__(self, o):
        return self.b._bin(N_SUB, self, self._lift(o))

    def __rsub__(self, o):
        return self.b._bin(N_SUB, self._lift(o), self)

    def __mul__(self, o):
        return self.b._bin(N_MUL, self, self._lift(o))

    __rmul__ = __mul__

    def __neg__(self):
        return self.b._un(N_NEG, self)


def _lin_add(a, b, sign):
    if a is None or b is None:
        return None
    c = (a[0] + sign * b[0]) % P
    t = dict(a[1])
    for k, w in b[1].items():
        t[k] = (t.get(k, 0) + sign * w) % P
    return (c, t)


def _lin_mul(a, b):
    if a is None or b is None:
        return None
    if not a[1]:
        return ((a[0] * b[0]) % P, {k: (w * a[0]) % P for k, w in b[1].items()})
    if not b[1]:
        return ((a[0] * b[0]) % P, {k: (w * b[0]) % P for k, w in a[1].items()})
    return None


class AirBuilder:
    """Records constraints of one chip as a hash-consed base-field expression DAG."""

    def __init__(self, prep_width: int, main_width: int):
        self.prep_width, self.main_width = prep_width, main_width
        self.nodes: list[tuple[int, int, int]] = []
        self._memo: dict[tuple[int, int, int], Expr] = {}
        self.constraints: list[int] = []
        self.sends: list[dict] = []
        self.receives: list[dict] = []
        self._cond: Expr | None = None

    # -- leaves ------------------------------------------------------------------------------
    def _node(self, op, a, b, deg, lin):
        key = (op, a, b)
        e = self._memo.get(key)
        if e is None:
            self.nodes.append(key)
            e = Expr(self, len(self.nodes) - 1, deg, lin)
            self._memo[key] = e
        return e

    def const(self, v: int) -> Expr:
        v = int(v) % P
        return self._node(N_CONST, v, 0, 0, (v, {}))

    def main(self, col: int, next: bool = False) -> Expr:
        assert 0 <= col < self.main_width
        lin = None if next else (0, {(1, col): 1})
        return self._node(N_MAIN, col, int(next), 1, lin)

    def prep(self, col: int, next: bool = False) -> Expr:
        assert 0 <= col < self.prep_width
        lin = None if next else (0, {(0, col): 1})
        return self._node(N_PREP, col, int(next), 1, lin)

    def pub(self, i: int) -> Expr:
        return self._node(N_PUB, i, 0, 0, None)

    def is_first_row(self) -> Expr:
        return self._node(N_IS_FIRST, 0, 0, 1, None)

    def is_last_row(self) -> Expr:
        return self._node(N_IS_LAST, 0, 0, 1, None)

    def is_transition(self) -> Expr:
        return self._node(N_IS_TRANS, 0, 0, 0, None)

    def _bin(self, op, x: Expr, y: Expr) -> Expr:
        if op == N_MUL:
            return self._node(op, x.id, y.id, x.deg + y.deg, _lin_mul(x.lin, y.lin))
        return self._node(op, x.id, y.id, max(x.deg, y.deg), _lin_add(x.lin, y.lin, 1 if op == N_ADD else -1))

    def _un(self, op, x: Expr) -> Expr:
        return self._node(op, x.id, 0, x.deg, _lin_add((0, {}), x.lin, -1))

    # -- constraints -------------------------------------------------------------------------
    def when(self, cond: Expr) -> "AirBuilder":
        sub = _Filtered(self, cond if self._cond is None else self._cond * cond)
        return sub

    def when_first_row(self):
        return self.when(self.is_first_row())

    def when_last_row(self):
        return self.when(self.is_last_row())

    def when_transition(self):
        return self.when(self.is_transition())

    def assert_zero(self, e):
        e = e if isinstance(e, Expr) else self.const(e)
        self.constraints.append(e.id)

    def assert_eq(self, a, b):
        a = a if isinstance(a, Expr) else self.const(a)
        self.assert_zero(a - b)

    # -- lookups -----------------------------------------------------------------------------
    def _vpc(self, e):
        e = e if isinstance(e, Expr) else self.const(e)
        if e.lin is None:
            raise ValueError("lookup values/multiplicities must be affine in the local row (VirtualPairCol)")
        c, terms = e.lin
        return (c, [(k[0], k[1], w) for k, w in sorted(terms.items()) if w])

    def send(self, kind: int, values, multiplicity, scope: int = SCOPE_LOCAL):
        self.sends.append(dict(kind=kind, scope=scope, values=[self._vpc(v) for v in values], mult=self._vpc(multiplicity)))

    def receive(self, kind: int, values, multiplicity, scope: int = SCOPE_LOCAL):
        self.receives.append(dict(kind=kind, scope=scope, values=[self._vpc(v) for v in values], mult=self._vpc(multiplicity)))

    def max_degree(self) -> int:
        nodes_deg = {}
        # degrees are stored on Expr objects; recover through the memo table
        for e in self._memo.values():
            nodes_deg[e.id] = e.deg
        return max([nodes_deg[c] for c in self.constraints], default=0)


class _Filtered:
    """`builder.when(cond)`: every assertion is multiplied by the condition (p3 FilteredAirBuilder)."""

    def __init__(self, base: AirBuilder, cond: Expr):
        self._base, self._c = base, cond

    def when(self, cond):
        return _Filtered(self._base, self._c * cond)

    def when_transition(self):
        return self.when(self._base.is_transition())

    def when_first_row(self):
        return self.when(self._base.is_first_row())

    def when_last_row(self):
        return self.when(self._base.is_last_row())

    def assert_zero(self, e):
        e = e if isinstance(e, Expr) else self._base.const(e)
        self._base.assert_zero(self._c * e)

    def assert_eq(self, a, b):
        a = a if isinstance(a, Expr) else self._base.const(a)
        self.assert_zero(a - b)


def _log2_ceil(x: int) -> int:
    return max(0, (x - 1).bit_length())


class Chip:
    """One table of the machine (the data a `Chip<F, A>` holds, crates/stark/src/chip.rs:19-34)."""

    def __init__(self, name: str, prep_width: int, main_width: int, eval_fn, *, local_only: bool = False,
                 global_scope: bool = False):
        self.name, self.prep_width, self.main_width = name, prep_width, main_width
        self.local_only, self.global_scope = local_only, global_scope
        b = AirBuilder(prep_width, main_width)
        eval_fn(b)
        self.builder = b
        deg = b.max_degree()
        if b.sends or b.receives:
            deg = max(deg, 3)
        self.log_quotient_degree = _log2_ceil(max(deg, 1) - 1) if deg > 1 else 0
        if self.global_scope:
            assert main_width >= 14

    # permutation.rs:18-23 with batch size 2^lqd
    @property
    def num_local_lookups(self) -> int:
        return sum(1 for l in self.builder.sends + self.builder.receives if l["scope"] == SCOPE_LOCAL)

    @property
    def perm_width_ef(self) -> int:
        n, bsz = self.num_local_lookups, 1 << self.log_quotient_degree
        return 0 if n == 0 else -(-n // bsz) + 1

    @property
    def num_constraints(self) -> int:
        c = len(self.builder.constraints)
        if self.perm_width_ef:
            c += self.perm_width_ef - 1 + 3
        if self.global_scope:
            c += 14
        return c

    @property
    def cost(self) -> int:  # chip.rs:154-163
        return self.prep_width + self.main_width + 4 * self.perm_width_ef + 4 * (1 << self.log_quotient_degree)

    def _words(self) -> list[int]:
        b = self.builder
        w = _str_words(self.name)
        w += [self.prep_width, self.main_width, self.log_quotient_degree, int(self.local_only), int(self.global_scope)]
        w += [len(b.sends), len(b.receives), len(b.nodes), len(b.constraints)]
        for l in b.sends + b.receives:
            w += [l["kind"], l["scope"], len(l["values"])]
            for c, terms in [l["mult"]] + l["values"]:
                w += [c, len(terms)]
                for is_main, col, wt in terms:
                    w += [is_main, col, wt]
        for op, a, bb in b.nodes:
            w += [op, a, bb]
        w += list(b.constraints)
        return w


def _str_words(s: str) -> list[int]:
    raw = s.encode()
    out = [len(raw)]
    raw = raw + b"\0" * (-len(raw) % 4)
    out += [int.from_bytes(raw[i:i + 4], "little") for i in range(0, len(raw), 4)]
    return out


class Machine:
    """A `StarkMachine`'s static description (crates/stark/src/machine.rs:38-75) plus the FRI
    parameters of `KoalaBearPoseidon2` (crates/stark/src/kb31_poseidon2.rs:203-213)."""

    MAGIC = 0x444D4B5A

    def __init__(self, chips: list[Chip], num_pv_elts: int, *, log_blowup: int = 1, num_queries: int = 84,
                 pow_bits: int = 16):
        self.chips, self.num_pv_elts = chips, num_pv_elts
        self.log_blowup, self.num_queries, self.pow_bits = log_blowup, num_queries, pow_bits
        for c in chips:
            if c.log_quotient_degree > log_blowup:
                raise ValueError(f"chip {c.name}: log_quotient_degree {c.log_quotient_degree} > log_blowup {log_blowup}")

    def chip(self, name: str) -> Chip:
        return next(c for c in self.chips if c.name == name)

    def descriptor(self) -> np.ndarray:
        w = [self.MAGIC, 1, len(self.chips), self.num_pv_elts, self.log_blowup, self.num_queries, self.pow_bits]
        for c in self.chips:
            w += c._words()
        return np.asarray(w, dtype=np.uint32)

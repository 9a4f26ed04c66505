"""Host side of GPU trace generation for the core ALU and control-flow chips: the chip
table, the `AluEvent` record layout and the padded height rule, mirroring what the Rust host keeps
when it hands a record's event vectors to `zkb200_generate_alu_trace` (include/zkb200.h).

Reference: `AluEvent` crates/core/executor/src/events/instr.rs:11-26 (#[repr(C)]: seven 32-bit
words, the #[repr(u8)] opcode in the low byte of the third); `next_power_of_two`
crates/core/machine/src/utils/mod.rs:101-125; chips crates/core/machine/src/alu/*/mod.rs."""
from __future__ import annotations

import numpy as np

from . import field as kb

# MachineAir::name -> (NUM_*_COLS, opcodes the chip receives, crates/core/executor/src/opcode.rs:25-49)
OPCODES = {"ADD": 0, "SUB": 1, "SLL": 9, "SRL": 10, "SRA": 11, "ROR": 12, "SLT": 13, "SLTU": 14, "AND": 15, "OR": 16,
           "XOR": 17, "NOR": 18, "CLZ": 19, "CLO": 20, "BEQ": 21, "BGEZ": 22, "BGTZ": 23, "BLEZ": 24, "BLTZ": 25, "BNE": 26,
           "Jump": 27, "Jumpi": 28, "JumpDirect": 29, "MEQ": 50, "MNE": 51, "WSBH": 52}
ALU_CHIPS = {
    "AddSub": (19, ("ADD", "SUB")),
    "Bitwise": (18, ("AND", "OR", "XOR", "NOR")),
    "Lt": (32, ("SLT", "SLTU")),
    "ShiftLeft": (44, ("SLL",)),
    "ShiftRight": (67, ("SRL", "SRA", "ROR")),
    "CloClz": (17, ("CLZ", "CLO")),
    # control flow: BranchEvent / JumpEvent records {pc, next_pc, next_next_pc, opcode, a, b, c}
    # (crates/core/executor/src/events/instr.rs:160-217), chips crates/core/machine/src/control_flow/
    "Branch": (62, ("BEQ", "BNE", "BLTZ", "BLEZ", "BGTZ", "BGEZ")),
    "Jump": (66, ("Jump", "Jumpi", "JumpDirect")),
    # MovCondEvent {pc, next_pc, opcode, a, b, c, prev_a} (instr.rs:287-302), crates/core/machine/src/misc/mov_cond/
    "MovCond": (32, ("MEQ", "MNE", "WSBH")),
}
# chips whose events are CompAluEvent records of sixteen words {shard, clk, pc, next_pc, opcode, hi, a, b, c, hi_record{value,
# shard, timestamp, prev_value, prev_shard, prev_timestamp}, hi_record_is_real} (crates/core/executor/src/events/instr.rs:47-73)
MEM_OPCODES = ("LB", "LBU", "LH", "LHU", "LW", "LWL", "LWR", "LL", "SB", "SH", "SW", "SWL", "SWR", "SC")
COMP_CHIPS = {"Mul": (58, ("MUL", "MULT", "MULTU")), "MemoryInstrs": (79, MEM_OPCODES)}
COMP_EVENT_WORDS = 16
OPCODES.update({"MUL": 2, "MULT": 3, "MULTU": 4})
OPCODES.update({name: 31 + i for i, name in enumerate(MEM_OPCODES)})
FLOW_CHIPS = ("Branch", "Jump")
EVENT_WORDS = 7          # AluEvent: pc, next_pc, opcode, hi, a, b, c;  Branch/JumpEvent: pc, next_pc, next_next_pc, opcode, a, b, c;
                         # MovCondEvent: pc, next_pc, opcode, a, b, c, prev_a
EVENT_BYTES = 28


CPU_WIDTH, CPU_EVENT_WORDS = 67, 28            # zkb200_cpu_event: the flattened CpuEvent + Instruction
MISC_WIDTH, MISC_EVENT_WORDS = 72, 15          # MiscEvent
PACKED_CHIPS = {"MemoryLocal": (56, 4)}       # width, events per row (seven-word MemoryLocalEvent records)
GLOBAL_WIDTH, GLOBAL_EVENT_WORDS = 99, 8        # GlobalLookupEvent: message[7], is_receive | kind << 8
SYSCALL_EVENT_WORDS = 14                       # SyscallEvent
MEMGLOBAL_EVENT_WORDS = 6                      # zkb200_memory_global_event: MemoryInitializeFinalizeEvent + prev_addr + position
# chip -> (width, event words): DivRem takes CompAluEvent records like Mul, the three syscall tables SyscallEvent records
ROW_CHIPS = {"DivRem": (106, 16), "SyscallCore": (11, 14), "SyscallPrecompile": (11, 14), "SyscallInstrs": (77, 14),
             "MemoryGlobalInit": (111, 6), "MemoryGlobalFinalize": (111, 6)}


def width(chip: str) -> int:
    if chip in PACKED_CHIPS:
        return PACKED_CHIPS[chip][0]
    if chip == "Cpu":
        return CPU_WIDTH
    if chip == "MiscInstrs":
        return MISC_WIDTH
    if chip in ROW_CHIPS:
        return ROW_CHIPS[chip][0]
    if chip == "Global":
        return GLOBAL_WIDTH
    return (ALU_CHIPS.get(chip) or COMP_CHIPS[chip])[0]


def events_per_row(chip: str) -> int:
    return PACKED_CHIPS[chip][1] if chip in PACKED_CHIPS else 1


def event_words(chip: str) -> int:
    if chip == "Cpu":
        return CPU_EVENT_WORDS
    if chip == "MiscInstrs":
        return MISC_EVENT_WORDS
    if chip in ROW_CHIPS:
        return ROW_CHIPS[chip][1]
    if chip == "Global":
        return GLOBAL_EVENT_WORDS
    return COMP_EVENT_WORDS if chip in COMP_CHIPS else EVENT_WORDS


def padded_log_height(n_events: int, fixed_log2_rows: int | None = None, chip: str | None = None) -> int:
    """`next_power_of_two(n, fixed_log2_rows)`: at least 16 rows, or the shape's fixed height.  `chip`: a chip that packs
    several events into a row (MemoryLocal) needs ceil(n / events_per_row) rows."""
    if chip is not None:
        n_events = -(-n_events // events_per_row(chip))
    if fixed_log2_rows is not None:
        if n_events > (1 << fixed_log2_rows):
            raise ValueError(f"fixed log2 rows is too small: got {n_events}, expected {1 << fixed_log2_rows}")
        return fixed_log2_rows
    return max(4, int(n_events - 1).bit_length() if n_events > 1 else 0)


def _alu_result(op: np.ndarray, b: np.ndarray, c: np.ndarray) -> np.ndarray:
    """`a` of well-formed events (what the MIPS executor would have recorded)."""
    b64, c64 = b.astype(np.uint64), c.astype(np.uint64)
    sh = c64 & np.uint64(31)
    sb, sc = b.astype(np.int32), c.astype(np.int32)
    out = np.zeros_like(b64)
    m32 = np.uint64(0xFFFFFFFF)

    def put(name, val):
        sel = op == OPCODES[name]
        out[sel] = val[sel] & m32

    put("ADD", b64 + c64)
    put("SUB", b64 - c64)
    put("SLL", b64 << sh)
    put("SRL", b64 >> sh)
    put("SRA", (sb.astype(np.int64) >> sh.astype(np.int64)).astype(np.uint64))
    put("ROR", (b64 >> sh) | (b64 << (np.uint64(32) - sh)))
    put("SLT", (sb < sc).astype(np.uint64))
    put("SLTU", (b < c).astype(np.uint64))
    put("AND", b64 & c64)
    put("OR", b64 | c64)
    put("XOR", b64 ^ c64)
    put("NOR", ~(b64 | c64))
    clz = np.array([32 - int(x).bit_length() for x in b], dtype=np.uint64) if b.size else np.zeros(0, np.uint64)
    clo = np.array([32 - int(x ^ 0xFFFFFFFF).bit_length() for x in b], dtype=np.uint64) if b.size else np.zeros(0, np.uint64)
    put("CLZ", clz)
    put("CLO", clo)
    return out.astype(np.uint32)


EDGE_OPERANDS = np.array([0, 1, 7, 8, 9, 31, 32, 33, 0x7F, 0x80, 0xFF, 0x100, 0x7FFF, 0x8000, 0xFFFF, 0x10000, 0x7FFFFF, 0x800000,
                          0xFFFFFF, 0x1000000, 0x7FFFFFFF, 0x80000000, 0x80000001, 0xFFFFFFFE, 0xFFFFFFFF], dtype=np.uint32)


def _flow_events(chip: str, n: int, rng, edges: bool) -> np.ndarray:
    """Well-formed BranchEvent / JumpEvent records: every program counter is a KoalaBear word (< p, which the
    chips range-check); next_next_pc is where the delay-slot semantics of the executor lands: the branch
    target next_pc + c when taken, else next_pc + 4; the register value b for Jump / Jumpi, next_pc + b for
    JumpDirect; a jump's a is the link value next_pc + 4."""
    ev = np.zeros((n, EVENT_WORDS), np.uint32)
    ev[:, 0] = rng.integers(1 << 22, kb.P - (1 << 22), n) & ~np.uint32(3)
    ev[:, 1] = ev[:, 0] + 4
    op = rng.choice([OPCODES[o] for o in ALU_CHIPS[chip][1]], n).astype(np.uint32)
    ev[:, 3] = op
    a = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    c = (rng.integers(-(1 << 17), 1 << 17, n) * 4).astype(np.int64).astype(np.uint32)
    if chip == "Branch":
        if edges:
            m = len(EDGE_OPERANDS)
            k = min(n, m * m)
            idx = np.arange(k)
            a[:k], b[:k] = EDGE_OPERANDS[idx // m], EDGE_OPERANDS[idx % m]
            lo, hi = k, min(n, k + 256)
            b[lo:hi] = a[lo:hi]
        sa, sb = a.astype(np.int32), b.astype(np.int32)
        eq, lt, gt = a == b, sa < sb, sa > sb
        taken = np.select([op == OPCODES["BEQ"], op == OPCODES["BNE"], op == OPCODES["BLTZ"], op == OPCODES["BLEZ"],
                           op == OPCODES["BGTZ"], op == OPCODES["BGEZ"]], [eq, ~eq, lt, lt | eq, gt, eq | gt])
        ev[:, 2] = np.where(taken, ev[:, 1] + c, ev[:, 1] + 4)
    else:
        a = (ev[:, 1] + 4).astype(np.uint32)          # link register value
        direct = op == OPCODES["JumpDirect"]
        target = (rng.integers(1 << 22, kb.P - (1 << 22), n) & ~np.uint32(3)).astype(np.uint32)
        b = np.where(direct, c, target).astype(np.uint32)        # BAL: pc-relative offset; J / JR: absolute target
        ev[:, 2] = np.where(direct, ev[:, 1] + c, target)
        if edges and n:
            # the largest KoalaBear word and its neighbours as jump targets (top byte 0x7f: the range checker's edge)
            k = min(n, 4)
            tops = np.array([kb.P - 1, 0x7F000000 - 4, 0x7E000000, 0x7EFFFFFC], np.uint32)[:k]
            ev[:k, 3] = OPCODES["Jump"]
            b[:k] = tops
            ev[:k, 2] = tops
    ev[:, 4], ev[:, 5], ev[:, 6] = a, b, c
    return ev


def _mov_cond_events(n: int, rng, edges: bool) -> np.ndarray:
    """MovCondEvent records {pc, next_pc, opcode, a, b, c, prev_a}: MEQ / MNE move b into a when c is / is not
    zero (else a keeps prev_a), WSBH swaps the bytes of each half word of b."""
    ev = np.zeros((n, EVENT_WORDS), np.uint32)
    ev[:, 0] = rng.integers(0, kb.P - 16, n) & ~np.uint32(3)
    ev[:, 1] = ev[:, 0] + 4
    op = rng.choice([OPCODES[o] for o in ALU_CHIPS["MovCond"][1]], n).astype(np.uint32)
    b = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    c = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    prev_a = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    # zero bytes in every position of c (the per-byte is-zero columns), whole-zero and half-zero words
    masks = np.array([0xFFFFFFFF, 0, 0xFFFF0000, 0x0000FFFF, 0xFF00FF00, 0x00FF00FF, 0xFFFFFF00, 0x00FFFFFF, 0xFF000000, 0x000000FF],
                     dtype=np.uint32)
    c &= masks[rng.integers(0, len(masks), n)]
    if edges and n:
        k = min(n, len(EDGE_OPERANDS))
        c[:k] = EDGE_OPERANDS[:k]
    prev_a[op == OPCODES["WSBH"]] = 0          # WSBH has no previous value of a (mov_cond/mod.rs: assert_word_zero)
    swapped = ((b & np.uint32(0x00FF00FF)) << 8) | ((b >> 8) & np.uint32(0x00FF00FF))
    a = np.select([op == OPCODES["MEQ"], op == OPCODES["MNE"]], [np.where(c == 0, b, prev_a), np.where(c != 0, b, prev_a)], swapped)
    ev[:, 2], ev[:, 3], ev[:, 4], ev[:, 5], ev[:, 6] = op, a.astype(np.uint32), b, c, prev_a
    return ev


def synthetic_events(chip: str, n: int, seed: int = 0, edges: bool = True) -> np.ndarray:
    """n well-formed events of the chip as (n, 7) uint32 words: seeded uniform operands, the first rows
    replaced by every pair of edge operands (byte boundaries, sign bits, equal and near-equal words)."""
    rng = np.random.default_rng(0xA1E00 + seed)
    if chip in FLOW_CHIPS:
        return _flow_events(chip, n, rng, edges)
    if chip == "MovCond":
        return _mov_cond_events(n, rng, edges)
    ev = np.zeros((n, EVENT_WORDS), np.uint32)
    ev[:, 0] = rng.integers(0, kb.P, n) & ~np.uint32(3)
    ev[:, 1] = ev[:, 0] + 4
    ev[:, 2] = rng.choice([OPCODES[o] for o in ALU_CHIPS[chip][1]], n)
    ev[:, 3] = 0
    ev[:, 5] = rng.integers(0, 1 << 32, n, dtype=np.uint64)
    ev[:, 6] = rng.integers(0, 1 << 32, n, dtype=np.uint64)
    if edges:
        m = len(EDGE_OPERANDS)
        k = min(n, m * m)
        idx = np.arange(k)
        ev[:k, 5] = EDGE_OPERANDS[idx // m]
        ev[:k, 6] = EDGE_OPERANDS[idx % m]
        lo, hi = k, min(n, k + 256)
        ev[lo:hi, 6] = ev[lo:hi, 5]                                                  # equal operands
        lo2, hi2 = hi, min(n, hi + 256)
        ev[lo2:hi2, 6] = ev[lo2:hi2, 5] ^ (np.uint32(1) << rng.integers(0, 32, hi2 - lo2).astype(np.uint32))   # one bit apart
    ev[:, 4] = _alu_result(ev[:, 2], ev[:, 5], ev[:, 6])
    return ev


def synthetic_mul_events(n: int, seed: int = 0, edges: bool = True) -> np.ndarray:
    """n well-formed CompAluEvent records of the Mul chip as (n, 16) uint32 words: MUL keeps the low word of b * c, MULT /
    MULTU write the high word to HI (hi_record: a memory write at clk + 4 of this shard, previous access earlier in this
    shard or in an earlier one)."""
    rng = np.random.default_rng(0x3A1 + seed)
    ev = np.zeros((n, COMP_EVENT_WORDS), np.uint32)
    shard = 3
    ev[:, 0] = shard
    ev[:, 1] = (5 + 8 * np.arange(1, n + 1)).astype(np.uint32)
    ev[:, 2] = rng.integers(0, kb.P - 16, n) & ~np.uint32(3)
    ev[:, 3] = ev[:, 2] + 4
    op = rng.choice([OPCODES[o] for o in COMP_CHIPS["Mul"][1]], n).astype(np.uint32)
    b = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    c = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    if edges:
        m = len(EDGE_OPERANDS)
        k = min(n, m * m)
        idx = np.arange(k)
        b[:k], c[:k] = EDGE_OPERANDS[idx // m], EDGE_OPERANDS[idx % m]
    signed = op == OPCODES["MULT"]
    prod_u = b.astype(np.uint64) * c.astype(np.uint64)
    prod_s = (b.astype(np.int32).astype(np.int64) * c.astype(np.int32).astype(np.int64)).astype(np.uint64)
    prod = np.where(signed, prod_s, prod_u)
    has_hi = op != OPCODES["MUL"]
    ev[:, 4] = op
    ev[:, 5] = np.where(has_hi, (prod >> np.uint64(32)).astype(np.uint32), 0)
    ev[:, 6] = (prod & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    ev[:, 7], ev[:, 8] = b, c
    earlier = rng.integers(0, 6, n) == 0
    ev[:, 9] = ev[:, 5]
    ev[:, 10] = shard
    ev[:, 11] = ev[:, 1] + 4
    ev[:, 12] = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    ev[:, 13] = np.where(earlier, rng.integers(1, shard, n), shard)
    ev[:, 14] = np.where(earlier, rng.integers(0, 1 << 22, n), ev[:, 11] - rng.integers(1, 9, n)).astype(np.uint32)
    ev[:, 15] = has_hi
    ev[~has_hi, 9:15] = 0
    return ev


def _sext(v: np.ndarray, bits: int) -> np.ndarray:
    v = v.astype(np.uint32)
    sign = (v >> np.uint32(bits - 1)) & np.uint32(1)
    return np.where(sign == 1, v | np.uint32((0xFFFFFFFF << bits) & 0xFFFFFFFF), v).astype(np.uint32)


def synthetic_mem_instr_events(n: int, seed: int = 0, edges: bool = True) -> np.ndarray:
    """n well-formed MemInstrEvent records of the MemoryInstrs chip as (n, 16) uint32 words, the #[repr(C)] image of
    crates/core/executor/src/events/instr.rs:114-136: {shard, clk, pc, next_pc, opcode, a, b, c, tag, record[6], prev_a_val}.
    Loads carry a MemoryReadRecord (tag 0: value, shard, timestamp, prev_shard, prev_timestamp, unused), stores a
    MemoryWriteRecord (tag 1: value, shard, timestamp, prev_value, prev_shard, prev_timestamp); `a` and the stored word follow
    execute_load / execute_store (crates/core/executor/src/executor.rs:1925-2090): halfword accesses are two-aligned, word
    accesses (LW, LL, SW, SC) four-aligned, the address stays below the field modulus."""
    rng = np.random.default_rng(0x4D31 + seed)
    ev = np.zeros((n, COMP_EVENT_WORDS), np.uint32)
    if n == 0:
        return ev
    shard = 3
    O = OPCODES
    op = rng.choice([O[o] for o in MEM_OPCODES], n).astype(np.uint32)
    if edges:
        op[: min(n, 4 * len(MEM_OPCODES))] = np.repeat([O[o] for o in MEM_OPCODES], 4)[: min(n, 4 * len(MEM_OPCODES))]
    addr = rng.integers(0, 0x7F000000, n).astype(np.uint32)
    small = rng.integers(0, 8, n) == 0                        # registers' address range: the upper three bytes are zero
    addr = np.where(small, addr & np.uint32(0xFF), addr)
    if edges:
        k = min(n, 4 * len(MEM_OPCODES))
        addr[:k] = (addr[:k] & ~np.uint32(3)) | (np.arange(k) % 4).astype(np.uint32)      # every opcode at every byte offset
    half = np.isin(op, [O["LH"], O["LHU"], O["SH"]])
    word = np.isin(op, [O["LW"], O["LL"], O["SW"], O["SC"]])
    addr = np.where(half, addr & ~np.uint32(1), np.where(word, addr & ~np.uint32(3), addr)).astype(np.uint32)
    c = _sext(rng.integers(0, 1 << 16, n).astype(np.uint32), 16)          # the sign-extended 16-bit offset
    b = (addr - c).astype(np.uint32)
    ls = addr & np.uint32(3)
    sh = (np.uint32(8) * ls).astype(np.uint32)
    mem = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)  # the aligned word before the instruction
    rt = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)   # register a before the instruction
    if edges:
        k = min(n, 4 * len(MEM_OPCODES))
        mem[:k:2] |= np.uint32(0x80808080)                                # negative bytes and halfwords
    full = np.uint32(0xFFFFFFFF)
    is_load = op <= O["LL"]
    byte = (mem >> sh) & np.uint32(0xFF)
    hw = (mem >> (np.uint32(8) * (ls & np.uint32(2)))) & np.uint32(0xFFFF)
    s24 = (np.uint32(24) - sh).astype(np.uint32)
    load_val = np.select(
        [op == O["LB"], op == O["LBU"], op == O["LH"], op == O["LHU"], op == O["LW"], op == O["LL"], op == O["LWL"], op == O["LWR"]],
        [_sext(byte, 8), byte, _sext(hw, 16), hw, mem, mem,
         (rt & ~(full << s24)) | (mem << s24), (rt & ~(full >> sh)) | (mem >> sh)], default=0).astype(np.uint32)
    hsh = (np.uint32(8) * (ls & np.uint32(2))).astype(np.uint32)
    store_val = np.select(
        [op == O["SB"], op == O["SH"], op == O["SW"], op == O["SC"], op == O["SWL"], op == O["SWR"]],
        [(mem & (full ^ (np.uint32(0xFF) << sh))) | ((rt & np.uint32(0xFF)) << sh),
         (mem & (full ^ (np.uint32(0xFFFF) << hsh))) | ((rt & np.uint32(0xFFFF)) << hsh),
         rt, rt, (mem & ~(full >> s24)) | (rt >> s24), (mem & ~(full << sh)) | (rt << sh)], default=0).astype(np.uint32)
    ev[:, 0] = shard
    ev[:, 1] = (5 + 8 * np.arange(1, n + 1)).astype(np.uint32)
    ev[:, 2] = rng.integers(0, kb.P - 16, n) & ~np.uint32(3)
    ev[:, 3] = ev[:, 2] + 4
    ev[:, 4] = op
    ev[:, 5] = np.where(is_load, load_val, np.where(op == O["SC"], 1, rt))
    ev[:, 6], ev[:, 7] = b, c
    ev[:, 8] = ~is_load
    earlier = rng.integers(0, 6, n) == 0
    prev_shard = np.where(earlier, rng.integers(1, shard, n), shard).astype(np.uint32)
    ts = ev[:, 1] + 1                                                     # MemoryAccessPosition::Memory
    prev_ts = np.where(earlier, rng.integers(0, 1 << 22, n), ts - rng.integers(1, 5, n)).astype(np.uint32)
    ev[:, 9] = np.where(is_load, mem, store_val)
    ev[:, 10] = shard
    ev[:, 11] = ts
    ev[:, 12] = np.where(is_load, prev_shard, mem)
    ev[:, 13] = np.where(is_load, prev_ts, prev_shard)
    ev[:, 14] = np.where(is_load, 0, prev_ts)
    ev[:, 15] = rt
    return ev


def synthetic_memory_local_events(n: int, seed: int = 0, shard: int = 3) -> np.ndarray:
    """n MemoryLocalEvent records as (n, 7) uint32 words {addr, initial {shard, timestamp, value}, final {shard, timestamp,
    value}} (crates/core/executor/src/events/memory.rs:228-237): distinct word-aligned addresses (registers included), the
    initial access in an earlier shard (or shard 0, clk 0 for untouched memory), the final one in this shard."""
    rng = np.random.default_rng(0x10CA1 + seed)
    ev = np.zeros((n, EVENT_WORDS), np.uint32)
    if n == 0:
        return ev
    addr = rng.choice(1 << 22, n, replace=False).astype(np.uint32) * np.uint32(4)
    addr[: min(n, 36)] = np.arange(min(n, 36), dtype=np.uint32)           # the register file
    untouched = rng.integers(0, 3, n) == 0
    ev[:, 0] = addr
    ev[:, 1] = np.where(untouched, 0, rng.integers(1, shard + 1, n))
    ev[:, 2] = np.where(untouched, 0, rng.integers(1, 1 << 22, n))
    ev[:, 3] = np.where(untouched, 0, rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32))
    ev[:, 4] = shard
    ev[:, 5] = rng.integers(1 << 22, 1 << 23, n)
    ev[:, 6] = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    return ev


ALL_OPCODES = {"ADD": 0, "SUB": 1, "MUL": 2, "MULT": 3, "MULTU": 4, "DIV": 5, "DIVU": 6, "MOD": 7, "MODU": 8, "SLL": 9, "SRL": 10,
               "SRA": 11, "ROR": 12, "SLT": 13, "SLTU": 14, "AND": 15, "OR": 16, "XOR": 17, "NOR": 18, "CLZ": 19, "CLO": 20,
               "BEQ": 21, "BGEZ": 22, "BGTZ": 23, "BLEZ": 24, "BLTZ": 25, "BNE": 26, "Jump": 27, "Jumpi": 28, "JumpDirect": 29,
               "SYSCALL": 30, "LB": 31, "LBU": 32, "LH": 33, "LHU": 34, "LW": 35, "LWL": 36, "LWR": 37, "LL": 38, "SB": 39,
               "SH": 40, "SW": 41, "SWL": 42, "SWR": 43, "SC": 44, "INS": 45, "MADDU": 46, "MSUBU": 47, "MADD": 48, "MSUB": 49,
               "MEQ": 50, "MNE": 51, "WSBH": 52, "EXT": 53, "TEQ": 54, "SEXT": 55}      # crates/core/executor/src/opcode.rs:25-89
# syscall codes the Cpu row looks at (crates/core/executor/src/syscalls/code.rs): HALT, SYS_EXT_GROUP, and others whose
# byte 3 is the number of extra cycles
_SYSCALL_CODES = np.array([0x00000000, 4246, 0x00000002, 0x00010005, 0x01010109, 0x00300130, 4003, 0x000000F0], np.uint32)


def synthetic_cpu_events(n: int, seed: int = 0, shard: int = 3) -> np.ndarray:
    """n `zkb200_cpu_event` records (include/zkb200.h) as (n, 28) uint32 words - what the shim writes per CpuEvent
    (crates/core/executor/src/events/cpu.rs:15-44) and its fetched Instruction: clk, pc, next_pc, next_next_pc, a, b, c, hi,
    flags, opcode | op_a << 8 | shard << 16, op_b, op_c, a_record[6], b_record[5], c_record[5].  Every opcode occurs; register
    a is written (most instructions), read (stores, branches) or untouched; b and c are register reads unless immediate;
    SYSCALL rows carry the syscall code as register a's previous value."""
    rng = np.random.default_rng(0xC9D + seed)
    ev = np.zeros((n, CPU_EVENT_WORDS), np.uint32)
    if n == 0:
        return ev
    O = ALL_OPCODES
    ops = np.array(list(O.values()), np.uint32)
    op = rng.choice(ops, n).astype(np.uint32)
    op[: min(n, len(ops))] = ops[: min(n, len(ops))]
    u32 = lambda: rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    clk = (5 + 5 * np.arange(1, n + 1) + (1 << 16) * (np.arange(n) % 3)).astype(np.uint32)       # exercises the 8-bit limb
    pc = (rng.integers(0, 1 << 22, n) * 4).astype(np.uint32)
    is_branch = (op >= O["BEQ"]) & (op <= O["BNE"])
    is_jump = (op >= O["Jump"]) & (op <= O["JumpDirect"])
    is_store = (op >= O["SB"]) & (op <= O["SWR"])
    is_sys = op == O["SYSCALL"]
    has_hi = np.isin(op, [O["MULT"], O["MULTU"], O["DIV"], O["DIVU"], O["MADDU"], O["MSUBU"], O["MADD"], O["MSUB"]])
    a, b, c = u32(), u32(), u32()
    imm_b = rng.integers(0, 4, n) == 0
    imm_c = (rng.integers(0, 2, n) == 0) | imm_b
    op_a = rng.integers(0, 34, n).astype(np.uint32)
    op_a[rng.integers(0, 8, n) == 0] = 0
    # register a: read for stores / branches, untouched for some jumps, written otherwise
    a_kind = np.where(is_store | is_branch, 1, np.where(is_jump & (rng.integers(0, 2, n) == 0), 0, 2)).astype(np.uint32)
    flags = (has_hi.astype(np.uint32) | (a_kind << 1) | ((~imm_b).astype(np.uint32) << 3) | ((~imm_c).astype(np.uint32) << 4)
             | (imm_b.astype(np.uint32) << 5) | (imm_c.astype(np.uint32) << 6))
    ev[:, 0], ev[:, 1], ev[:, 2] = clk, pc, pc + 4
    ev[:, 3] = np.where(is_branch | is_jump, (rng.integers(0, 1 << 22, n) * 4).astype(np.uint32), pc + 8)
    ev[:, 4], ev[:, 5], ev[:, 6] = a, b, c
    ev[:, 7] = np.where(has_hi, u32(), 0)
    ev[:, 8] = flags
    ev[:, 9] = op | (op_a << 8) | (np.uint32(shard) << 16)
    ev[:, 10] = np.where(imm_b, b, rng.integers(0, 34, n))
    ev[:, 11] = np.where(imm_c, c, rng.integers(0, 34, n))

    def prev(cur_ts):
        earlier = (rng.integers(0, 6, n) == 0) & (shard > 1)
        pshard = np.where(earlier, rng.integers(1, max(shard, 2), n), shard).astype(np.uint32)
        pts = np.where(earlier, rng.integers(0, 1 << 22, n), cur_ts.astype(np.int64) - rng.integers(1, 5, n)).astype(np.uint32)
        return pshard, pts

    # a_record at clk + 0 (MemoryAccessPosition::A... the position offsets only need to be distinct and increasing here)
    ps, pt = prev(clk + 3)
    prev_a = np.where(is_sys, rng.choice(_SYSCALL_CODES, n), u32()).astype(np.uint32)
    wr = a_kind == 2
    ev[:, 12], ev[:, 13], ev[:, 14] = a, shard, clk + 3
    ev[:, 15] = np.where(wr, prev_a, ps)
    ev[:, 16] = np.where(wr, ps, pt)
    ev[:, 17] = np.where(wr, pt, 0)
    ev[a_kind == 0, 12:18] = 0
    for col0, val, imm, off in ((18, b, imm_b, 1), (23, c, imm_c, 2)):
        ps, pt = prev(clk + off)
        ev[:, col0], ev[:, col0 + 1], ev[:, col0 + 2], ev[:, col0 + 3], ev[:, col0 + 4] = val, shard, clk + off, ps, pt
        ev[imm, col0:col0 + 5] = 0
    return ev


MISC_OPCODES = ("SEXT", "EXT", "INS", "MADDU", "MSUBU", "MADD", "MSUB", "TEQ")


def synthetic_misc_events(n: int, seed: int = 0, shard: int = 3) -> np.ndarray:
    """n well-formed MiscEvent records of the MiscInstrs chip as (n, 15) uint32 words {shard, clk, pc, next_pc, opcode, a, b,
    c, prev_a, hi_record{value, shard, timestamp, prev_value, prev_shard, prev_timestamp}}
    (crates/core/executor/src/events/instr.rs:241-261), `a` and the HI write following execute_sext / ext / ins / maddu /
    msubu / madd / msub / teq (crates/core/executor/src/executor.rs:1686-1828): SEXT with c = 0 (byte) and c > 0 (halfword),
    EXT with lsb + msbd <= 31, INS with lsb <= msb, TEQ with a != b except for shared low bytes."""
    rng = np.random.default_rng(0x315C + seed)
    ev = np.zeros((n, MISC_EVENT_WORDS), np.uint32)
    if n == 0:
        return ev
    O = ALL_OPCODES
    u32 = lambda: rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    op = rng.choice([O[o] for o in MISC_OPCODES], n).astype(np.uint32)
    op[: min(n, 8)] = [O[o] for o in MISC_OPCODES][: min(n, 8)]
    b, c, prev_a = u32(), u32(), u32()
    b[rng.integers(0, 4, n) == 0] |= np.uint32(0x8080)                     # negative bytes / halfwords
    a = np.zeros(n, np.uint32)
    hi_prev, hi_new = u32(), np.zeros(n, np.uint32)
    full = np.uint64(0xFFFFFFFFFFFFFFFF)
    # SEXT
    m = op == O["SEXT"]
    c[m] = rng.integers(0, 2, int(m.sum()))
    a[m] = np.where(c[m] > 0, _sext(b[m] & np.uint32(0xFFFF), 16), _sext(b[m] & np.uint32(0xFF), 8))
    # TEQ: a = src1, b = src2, differing, some bytes equal
    m = op == O["TEQ"]
    c[m] = 0
    flip = (np.uint32(0xFF) << (np.uint32(8) * rng.integers(0, 4, int(m.sum())).astype(np.uint32))).astype(np.uint32)
    a[m] = b[m] ^ np.where(rng.integers(0, 2, int(m.sum())) == 0, flip, rng.integers(1, 1 << 32, int(m.sum()), dtype=np.uint64).astype(np.uint32))
    # EXT: a = (b & mask(msbd + lsb + 1)) >> lsb
    m = op == O["EXT"]
    k = int(m.sum())
    lsb = rng.integers(0, 32, k)
    msbd = np.array([rng.integers(0, 32 - l) for l in lsb], np.int64) if k else np.zeros(0, np.int64)
    c[m] = (lsb | (msbd << 5)).astype(np.uint32)
    top = (msbd + lsb + 1).astype(np.uint64)
    mask = np.where(top == 32, np.uint64(0xFFFFFFFF), (np.uint64(1) << top) - np.uint64(1)).astype(np.uint32)
    a[m] = (b[m] & mask) >> lsb.astype(np.uint32)
    # INS: a = (prev_a & ~field) | ((b << lsb) & field)
    m = op == O["INS"]
    k = int(m.sum())
    lsb = rng.integers(0, 32, k)
    msb = np.array([rng.integers(l, 32) for l in lsb], np.int64) if k else np.zeros(0, np.int64)
    c[m] = (lsb | (msb << 5)).astype(np.uint32)
    wid = (msb - lsb + 1).astype(np.uint64)
    mask = np.where(wid == 32, np.uint64(0xFFFFFFFF), (np.uint64(1) << wid) - np.uint64(1)).astype(np.uint32)
    field = (mask.astype(np.uint64) << lsb.astype(np.uint64)).astype(np.uint32)
    a[m] = (prev_a[m] & ~field) | ((b[m].astype(np.uint64) << lsb.astype(np.uint64)).astype(np.uint32) & field)
    # MADDU / MSUBU / MADD / MSUB: (hi, lo) +- b * c
    for name, signed, add in (("MADDU", False, True), ("MSUBU", False, False), ("MADD", True, True), ("MSUB", True, False)):
        m = op == O[name]
        if not m.any():
            continue
        if signed:
            prod = (b[m].astype(np.int32).astype(np.int64) * c[m].astype(np.int32).astype(np.int64)).astype(np.uint64)
        else:
            prod = b[m].astype(np.uint64) * c[m].astype(np.uint64)
        addend = (hi_prev[m].astype(np.uint64) << np.uint64(32)) + prev_a[m].astype(np.uint64)
        out = (addend + prod) & full if add else (addend - prod) & full
        a[m] = (out & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        hi_new[m] = (out >> np.uint64(32)).astype(np.uint32)
    is_mac = (op >= O["MADDU"]) & (op <= O["MSUB"])
    ev[:, 0] = shard
    ev[:, 1] = (5 + 5 * np.arange(1, n + 1)).astype(np.uint32)
    ev[:, 2] = rng.integers(0, kb.P - 16, n) & ~np.uint32(3)
    ev[:, 3] = ev[:, 2] + 4
    ev[:, 4], ev[:, 5], ev[:, 6], ev[:, 7], ev[:, 8] = op, a, b, c, prev_a
    earlier = rng.integers(0, 6, n) == 0
    ts = ev[:, 1] + 4
    ev[:, 9], ev[:, 10], ev[:, 11], ev[:, 12] = hi_new, shard, ts, hi_prev
    ev[:, 13] = np.where(earlier, rng.integers(1, shard, n), shard)
    ev[:, 14] = np.where(earlier, rng.integers(0, 1 << 22, n), ts - rng.integers(1, 9, n)).astype(np.uint32)
    ev[~is_mac, 9:15] = 0
    return ev


def _write_record(rng, n: int, shard: int, ts: np.ndarray, value: np.ndarray, prev_value: np.ndarray) -> np.ndarray:
    """(n, 6) MemoryWriteRecord words {value, shard, timestamp, prev_value, prev_shard, prev_timestamp}, the previous access in
    the same shard at an earlier clock or in an earlier shard."""
    rec = np.zeros((n, 6), np.uint32)
    earlier = rng.integers(0, 6, n) == 0
    rec[:, 0], rec[:, 1], rec[:, 2], rec[:, 3] = value, shard, ts, prev_value
    rec[:, 4] = np.where(earlier, rng.integers(1, max(2, shard), n), shard)
    rec[:, 5] = np.where(earlier, rng.integers(0, 1 << 22, n), ts - rng.integers(1, 9, n)).astype(np.uint32)
    return rec


def synthetic_div_rem_events(n: int, seed: int = 0, shard: int = 3, edges: bool = True) -> np.ndarray:
    """n well-formed CompAluEvent records of the DivRem chip as (n, 16) uint32 words {shard, clk, pc, next_pc, opcode, hi, a, b,
    c, hi_record[6], hi_record_is_real} (crates/core/executor/src/events/instr.rs:47-73) for DIV / DIVU / MOD / MODU, lo / hi as
    get_quotient_and_remainder leaves them (crates/core/executor/src/utils.rs:33-43).  `edges`: also c = 0, INT_MIN / -1,
    c = +-1 and equal operands - the rows where the reference's C++ twin differs from its Rust are among them."""
    rng = np.random.default_rng(0xD17 + seed)
    ev = np.zeros((n, 16), np.uint32)
    if n == 0:
        return ev
    O = ALL_OPCODES
    ops = [O["DIV"], O["DIVU"], O["MOD"], O["MODU"]]
    op = rng.choice(ops, n).astype(np.uint32)
    op[: min(n, 4)] = ops[: min(n, 4)]
    b = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    c = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    small = rng.integers(0, 3, n) == 0
    c[small] = (c[small] >> rng.integers(8, 31, int(small.sum())).astype(np.uint32)) | np.uint32(1)
    neg = rng.integers(0, 4, n) == 0
    c[neg] = (~c[neg]) + np.uint32(1)
    if edges and n >= 16:
        k = np.arange(n)
        c[k % 23 == 5] = 0
        c[k % 29 == 7] = 0xFFFFFFFF
        b[k % 58 == 7] = 0x80000000                       # with c = -1: the overflow row
        c[k % 31 == 9] = 1
        m = k % 37 == 11
        c[m] = b[m]
        b[k % 41 == 13] = 0x80000000
    sgn = (op == O["DIV"]) | (op == O["MOD"])
    sb, sc = b.astype(np.int32).astype(np.int64), c.astype(np.int32).astype(np.int64)
    nz = c != 0
    q = np.full(n, 0xFFFFFFFF, np.uint64)
    r = b.astype(np.uint64)
    scs = np.where(nz, sc, 1)
    qs = np.abs(sb) // np.abs(scs) * np.sign(sb) * np.sign(scs)            # truncation toward zero
    rs = sb - qs * scs
    cu = np.where(nz, c, 1).astype(np.uint64)
    q = np.where(nz, np.where(sgn, qs.astype(np.uint64) & np.uint64(0xFFFFFFFF), b.astype(np.uint64) // cu), q)
    r = np.where(nz, np.where(sgn, rs.astype(np.uint64) & np.uint64(0xFFFFFFFF), b.astype(np.uint64) % cu), r)
    ev[:, 0] = shard
    ev[:, 1] = (5 + 5 * np.arange(1, n + 1)).astype(np.uint32)
    ev[:, 2] = rng.integers(0, kb.P - 16, n) & ~np.uint32(3)
    ev[:, 3] = ev[:, 2] + 4
    ev[:, 4], ev[:, 5], ev[:, 6], ev[:, 7], ev[:, 8] = op, r.astype(np.uint32), q.astype(np.uint32), b, c
    is_div = (op == O["DIV"]) | (op == O["DIVU"])
    rec = _write_record(rng, n, shard, ev[:, 1] + 4, r.astype(np.uint32), rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32))
    ev[:, 9:15] = np.where(is_div[:, None], rec, 0)
    ev[:, 15] = is_div
    return ev


# SyscallCode values (crates/core/executor/src/syscalls/code.rs:28-160): bytes 0-1 the id (byte 1 != 0 for the linux calls), byte 2
# "send to table" (1 for the precompiles), byte 3 the extra cycles: HALT, WRITE, ENTER_UNCONSTRAINED, EXIT_UNCONSTRAINED,
# COMMIT, COMMIT_DEFERRED_PROOFS, VERIFY_ZKM_PROOF, SYSHINTLEN, SYSHINTREAD, SYS_EXT_GROUP, SYS_READ, SYS_WRITE, SYS_BRK,
# SYS_MMAP2, SHA_EXTEND, SHA_COMPRESS, ED_DECOMPRESS, KECCAK_SPONGE, BLS12381_DECOMPRESS, UINT256_MUL
_SYSCALL_TABLE = np.array([0x00000000, 0x00000002, 0x00000003, 0x00000004, 0x00000010, 0x0000001A, 0x0000001B, 0x000000F0, 0x000000F1, 4246,
                           4003, 4004, 4045, 4090, 0x30010005, 0x01010006, 0x00010008, 0x01010009, 0x0001001C, 0x0101001D], np.uint32)


def syscall_public_values(seed: int = 0):
    """The public values SyscallInstrs' COMMIT / COMMIT_DEFERRED_PROOFS / HALT rows are checked against
    (crates/core/machine/src/syscall/instructions/air.rs:274-358): committed_value_digest (8 words), deferred_proofs_digest (8 field
    elements), exit_code."""
    rng = np.random.default_rng(0x9B + seed)
    return (rng.integers(0, 1 << 32, 8, dtype=np.uint64).astype(np.uint32), rng.integers(0, kb.P, 8).astype(np.uint32), 7)


def synthetic_syscall_events(n: int, seed: int = 0, shard: int = 3, kind: str = "instrs") -> np.ndarray:
    """n well-formed SyscallEvent records as (n, 14) uint32 words {pc, next_pc, shard, clk, a_record[6], a_record_is_real,
    syscall_id, arg1, arg2} (crates/core/executor/src/events/syscall.rs:8-29); a_record.prev_value holds the syscall code the
    instruction read from $v0, a_record.value what it left there.  kind "instrs": every syscall of the shard (SyscallInstrs), as
    the executor emits them (crates/core/executor/src/executor.rs execute_syscall) - $v0 unchanged except for
    ENTER_UNCONSTRAINED (0), SYSHINTLEN and the linux calls (a result), syscall_id EXIT_UNCONSTRAINED for ENTER_UNCONSTRAINED,
    next_pc = 0 and arg1 = the exit code for HALT / SYS_EXT_GROUP, COMMIT / COMMIT_DEFERRED_PROOFS with a digest index below 8
    and the digest word of syscall_public_values(seed), arguments below the KoalaBear modulus on both sides of its top byte;
    "core": only the events SyscallCore keeps (prev_value byte 2 = 1 or byte 1 != 0, chip.rs:233-240); "precompile": one event
    per precompile event with prev_value = 1 / value = v0 for the Linux ones and prev_value = 0 otherwise (include/syscall.hpp
    precompile_event_to_row)."""
    rng = np.random.default_rng(0x5C11 + seed)
    ev = np.zeros((n, SYSCALL_EVENT_WORDS), np.uint32)
    if n == 0:
        return ev
    u32 = lambda: rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    table = _SYSCALL_TABLE
    if kind == "core":
        table = table[(((table >> 16) & 0xFF) == 1) | (((table >> 8) & 0xFF) != 0)]
    code = rng.choice(table, n).astype(np.uint32)
    code[: min(n, len(table))] = table[: min(n, len(table))]

    def below_p():
        v = u32() & np.uint32(0x7EFFFFFF)
        pick = rng.integers(0, 6, n)
        v = np.where(pick == 0, np.uint32(0x7F000000), v)                       # the largest top byte: the low bytes must be zero
        return np.where(pick == 1, (v & np.uint32(0x00FFFFFF)) | np.uint32(0x7E000000), v).astype(np.uint32)
    arg1, arg2 = below_p(), below_p()
    sid = code & np.uint32(0xFFFF)
    digest, deferred, exit_code = syscall_public_values(seed)
    idx = rng.integers(0, 8, n)
    is_commit, is_deferred = sid == 0x10, sid == 0x1A
    arg1 = np.where(is_commit | is_deferred, idx, arg1).astype(np.uint32)
    arg2 = np.where(is_commit, digest[idx], np.where(is_deferred, deferred[idx], arg2)).astype(np.uint32)
    is_halt = (sid == 0) | (sid == 4246)
    arg1 = np.where(is_halt, exit_code, arg1).astype(np.uint32)
    linux = ((code >> 8) & 0xFF) != 0
    value = np.where(sid == 0x03, 0, np.where((sid == 0xF0) | linux, u32(), code)).astype(np.uint32)
    prev_value = code.copy()
    if kind == "precompile":
        is_linux = rng.integers(0, 3, n) == 0
        prev_value = is_linux.astype(np.uint32)
        value = np.where(is_linux, u32(), 0).astype(np.uint32)
    ev[:, 0] = rng.integers(0, kb.P - 16, n) & ~np.uint32(3)
    ev[:, 1] = np.where(is_halt & (kind != "precompile"), 0, ev[:, 0] + 4)
    ev[:, 2] = shard
    ev[:, 3] = (5 + 5 * np.arange(1, n + 1)).astype(np.uint32)
    ev[:, 4:10] = _write_record(rng, n, shard, ev[:, 3], value, prev_value)
    if kind == "precompile":
        ev[:, 5:7] = 0
        ev[:, 8:10] = 0                                   # a default MemoryWriteRecord apart from the two fields above
    ev[:, 10] = 0 if kind == "precompile" else 1
    ev[:, 11] = np.where(sid == 0x03, 0x04, sid)
    ev[:, 12], ev[:, 13] = arg1, arg2
    return ev


def memory_global_records(events: np.ndarray, previous_addr: int) -> np.ndarray:
    """The flattened records `zkb200_generate_alu_trace("MemoryGlobalInit" | "MemoryGlobalFinalize")` takes: the (n, 4)
    MemoryInitializeFinalizeEvent records {addr, value, shard, timestamp} sorted by address as generate_trace does
    (crates/core/machine/src/memory/global.rs:130), each followed by the address it is compared with (the previous event's; for
    the first event the public values' previous_init_addr / previous_finalize_addr) and its position (bit 0 first, bit 1 last):
    what the reference's second, sequential loop (global.rs:150-180) reads from the neighbouring row."""
    ev = np.ascontiguousarray(events, dtype=np.uint32).reshape(-1, 4)
    ev = ev[np.argsort(ev[:, 0], kind="stable")]
    n = len(ev)
    out = np.zeros((n, MEMGLOBAL_EVENT_WORDS), np.uint32)
    if n == 0:
        return out
    out[:, :4] = ev
    out[0, 4] = previous_addr
    out[1:, 4] = ev[:-1, 0]
    out[0, 5] |= 1
    out[n - 1, 5] |= 2
    return out


def synthetic_memory_global_events(n: int, seed: int = 0, shard: int = 3) -> np.ndarray:
    """n MemoryInitializeFinalizeEvent records (n, 4) {addr, value, shard, timestamp} with distinct addresses below the
    KoalaBear modulus, UNSORTED as they sit in record.global_memory_initialize_events: neighbouring addresses that differ in
    their lowest and in their highest bits, addresses with the top byte 0x7E / 0x7F prefix bits set."""
    rng = np.random.default_rng(0x3E3 + seed)
    ev = np.zeros((n, 4), np.uint32)
    if n == 0:
        return ev
    addr = set()
    base = int(rng.integers(1, 1 << 20))
    while len(addr) < n:
        r = int(rng.integers(0, 4))
        if r == 0:
            base += int(rng.integers(1, 3))
        elif r == 1:
            base += int(rng.integers(1, 1 << 12))
        elif r == 2:
            base = int(rng.integers(1, kb.P))
        else:
            base = int(rng.choice([0x7E000000, 0x7F000000, 0x3F000000, 0x7C000000])) - int(rng.integers(1, 1 << 16))
        if 0 < base < kb.P:
            addr.add(base)
    a = np.array(sorted(addr), np.uint32)
    rng.shuffle(a)
    ev[:, 0] = a
    ev[:, 1] = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    ev[:, 2] = rng.integers(0, shard + 1, n)
    ev[:, 3] = rng.integers(0, 1 << 24, n)
    return ev


def synthetic_global_events(n: int, seed: int = 0, shard: int = 3) -> np.ndarray:
    """n GlobalLookupEvent records as (n, 8) uint32 words {message[7], is_receive | kind << 8}
    (crates/core/executor/src/events/global.rs:6-15): memory messages {shard, timestamp, addr, four value bytes} of kind
    LookupKind::Memory = 1 (crates/core/machine/src/memory/local.rs generate_dependencies), sent and received, and syscall
    messages {shard, clk, syscall_id, four half-words} of kinds Syscall = 6 / SyscallResult = 8 (crates/stark/src/lookup/lookup.rs:25-47,
    syscall/chip.rs:152-167);
    message[0] below 2^16 as the chip range-checks it."""
    rng = np.random.default_rng(0x610B + seed)
    ev = np.zeros((n, GLOBAL_EVENT_WORDS), np.uint32)
    if n == 0:
        return ev
    mem = rng.integers(0, 4, n) != 0
    ev[:, 0] = rng.integers(0, shard + 1, n)
    ev[:, 1] = rng.integers(0, 1 << 24, n)
    ev[:, 2] = np.where(mem, rng.integers(0, kb.P, n), rng.choice(_SYSCALL_TABLE & 0xFFFF, n))
    ev[:, 3:7] = np.where(mem[:, None], rng.integers(0, 256, (n, 4)), rng.integers(0, 1 << 16, (n, 4)))
    kind = np.where(mem, 1, rng.choice([6, 8], n)).astype(np.uint32)
    ev[:, 7] = rng.integers(0, 2, n).astype(np.uint32) | (kind << np.uint32(8))
    return ev


def memory_global_lookup_events(events: np.ndarray, finalize: bool) -> np.ndarray:
    """MemoryGlobalChip::generate_dependencies (crates/core/machine/src/memory/global.rs:62-97): one GlobalLookupEvent (n, 8)
    per address-sorted memory event - an initialisation sends {0, 0, addr, value bytes}, a finalisation receives {shard,
    timestamp, addr, value bytes} - of kind LookupKind::Memory."""
    ev = np.ascontiguousarray(events, dtype=np.uint32).reshape(-1, 4)
    ev = ev[np.argsort(ev[:, 0], kind="stable")]
    out = np.zeros((len(ev), GLOBAL_EVENT_WORDS), np.uint32)
    if finalize:
        out[:, 0], out[:, 1] = ev[:, 2], ev[:, 3]
    out[:, 2] = ev[:, 0]
    for k in range(4):
        out[:, 3 + k] = (ev[:, 1] >> np.uint32(8 * k)) & np.uint32(0xFF)
    out[:, 7] = np.uint32(int(finalize)) | np.uint32(1 << 8)
    return out


def syscall_global_lookup_events(events: np.ndarray, precompile: bool = False) -> np.ndarray:
    """SyscallChip::generate_dependencies (crates/core/machine/src/syscall/chip.rs:119-172): two GlobalLookupEvents (2n, 8) per
    SyscallEvent record of the table - {shard, clk, syscall_id, arg1 half-words, arg2 half-words} of kind Syscall = 6 and {shard,
    clk, syscall_id, result half-words, 0, 0} of kind SyscallResult = 8, the result only for the linux calls - sent by a core
    shard, received by a precompile shard."""
    ev = np.ascontiguousarray(events, dtype=np.uint32).reshape(-1, SYSCALL_EVENT_WORDS)
    n = len(ev)
    out = np.zeros((2 * n, GLOBAL_EVENT_WORDS), np.uint32)
    prev_value, value = ev[:, 7], ev[:, 4]
    linux = (prev_value == 1) if precompile else (((prev_value >> 8) & 0xFF) != 0)
    result = np.where(linux, value, 0).astype(np.uint32)
    for k in (0, 1):
        out[k::2, 0], out[k::2, 1], out[k::2, 2] = ev[:, 2], ev[:, 3], ev[:, 11]
    out[0::2, 3], out[0::2, 4] = ev[:, 12] & 0xFFFF, ev[:, 12] >> 16
    out[0::2, 5], out[0::2, 6] = ev[:, 13] & 0xFFFF, ev[:, 13] >> 16
    out[1::2, 3], out[1::2, 4] = result & 0xFFFF, result >> 16
    out[0::2, 7] = np.uint32(int(precompile)) | np.uint32(6 << 8)
    out[1::2, 7] = np.uint32(int(precompile)) | np.uint32(8 << 8)
    return out

"""The KeccakSponge precompile chip as data: its `Air::eval` restated with ziren_b200/air.py, and the tables that
balance its lookups in a synthetic shard (no guest ELF can be executed here).

Reference:
  * `impl Air for KeccakSpongeChip`, crates/core/machine/src/syscall/precompiles/keccak_sponge/air.rs:25-285
    (flags, memory accesses, state <-> permutation limbs, syscall receive, XOR lookups, absorbed-length rules),
    `eval_memory_access` / `eval_memory_access_timestamp` / `eval_range_check_24bits`,
    crates/core/machine/src/air/memory.rs:18-172; `XorOperation::eval`, crates/core/machine/src/operations/xor.rs:41-58;
    lookup tuples `send_byte` / `receive_syscall`, crates/stark/src/air/builder.rs:120-150, :356-380.
  * `p3_keccak_air::KeccakAir::eval` (air.rs:122-125 runs it over the first 2633 columns through SubAirBuilder) is in
    the UN-VENDORED Plonky3 (keccak-air/src/{air,round_flags,logic}.rs at the commit Cargo.lock pins); its published
    constraints are restated here: round flags, preimage, C' = xor3(C, C[x-1], rot C[x+1]), A limbs = bits of
    xor3(A', C, C'), the parity cubic on the column sums of A', A'' = B ^ andn(B[x+1], B[x+2]), the iota bits, and
    "this round's output is the next round's input".
  * receivers: ByteChip (crates/core/machine/src/bytes/air.rs:21-78: XOR, U8Range, U16Range of the ten opcodes),
    MemoryLocalChip's local lookups (crates/core/machine/src/memory/local.rs:208-275; its Global-kind sends are left
    out), SyscallChip's send (crates/core/machine/src/syscall/chip.rs).
The traces these constraints are checked on are the reference-identical rows of the row fillers
(ziren_b200/csrc/tracegen_keccak.cuh, oracle/tracegen_keccak.h)."""
from __future__ import annotations

import numpy as np

from . import keccak_sponge as ks
from .air import KIND_BYTE, KIND_MEMORY, KIND_SYSCALL, Chip

C = ks.Cols
RATE, STATE, OUT, ROUNDS = ks.RATE_U32S, ks.STATE_U32S, ks.OUTPUT_U32S, ks.NUM_ROUNDS
BYTE_XOR, BYTE_U8RANGE, BYTE_U16RANGE = 2, 4, 8          # ByteOpcode, crates/core/executor/src/opcode.rs:195-216
SYSCALL_ID = 0x01_01_00_09 & 0xFFFF                        # SyscallCode::KECCAK_SPONGE.syscall_id(), syscalls/code.rs:57, :269


def _xor(x, y):
    return x + y - 2 * (x * y)


def _xor3(x, y, z):
    return _xor(x, _xor(y, z))


def _andn(x, y):
    return (1 - x) * y


def _assert_bool(b, x):
    b.assert_zero(x * (x - 1))


def _eval_p3_keccak(b):
    """KeccakAir::eval over columns 0..2633."""
    m = b.main
    nxt = lambda c: b.main(c, next=True)
    flags = [m(C.STEP_FLAGS + i) for i in range(ROUNDS)]
    # round_flags.rs
    b.when_first_row().assert_eq(flags[0], 1)
    for i in range(1, ROUNDS):
        b.when_first_row().assert_zero(flags[i])
    for i in range(ROUNDS):
        b.when_transition().assert_eq(nxt(C.STEP_FLAGS + (i + 1) % ROUNDS), flags[i])
    first_step, final_step = flags[0], flags[ROUNDS - 1]
    not_final_step = 1 - final_step
    pre = lambda y, x, l: C.PREIMAGE + (y * 5 + x) * 4 + l
    a_col = lambda y, x, l: C.A + (y * 5 + x) * 4 + l
    c_ = lambda x, z: m(C.C + x * 64 + z)
    cp = lambda x, z: m(C.C_PRIME + x * 64 + z)
    ap = lambda y, x, z: m(C.A_PRIME + (y * 5 + x) * 64 + z)
    app = lambda y, x, l: m(C.A_PRIME_PRIME + (y * 5 + x) * 4 + l)
    bits00 = lambda z: m(C.A_PRIME_PRIME_0_0_BITS + z)
    appp00 = lambda l: m(C.A_PRIME_PRIME_PRIME_0_0_LIMBS + l)
    for y in range(5):
        for x in range(5):
            for l in range(4):
                b.when(first_step).assert_eq(m(pre(y, x, l)), m(a_col(y, x, l)))
    for y in range(5):
        for x in range(5):
            for l in range(4):
                b.when(not_final_step).when_transition().assert_eq(m(pre(y, x, l)), nxt(pre(y, x, l)))
    export = m(C.EXPORT)
    _assert_bool(b, export)
    b.when(not_final_step).assert_zero(export)
    for x in range(5):
        for z in range(64):
            _assert_bool(b, c_(x, z))
            b.assert_eq(cp(x, z), _xor3(c_(x, z), c_((x + 4) % 5, z), c_((x + 1) % 5, (z + 63) % 64)))
    for y in range(5):
        for x in range(5):
            for l in range(4):
                acc = b.const(0)
                for z in reversed(range(16 * l, 16 * l + 16)):
                    _assert_bool(b, ap(y, x, z))
                    acc = acc + acc + _xor3(ap(y, x, z), c_(x, z), cp(x, z))
                b.assert_eq(acc, m(a_col(y, x, l)))
    for x in range(5):
        for z in range(64):
            diff = ap(0, x, z) + ap(1, x, z) + ap(2, x, z) + ap(3, x, z) + ap(4, x, z) - cp(x, z)
            b.assert_zero(diff * (diff - 2) * (diff - 4))

    def bb(x, y, z):       # columns.rs b(): an alias of A'
        a, bcol = (x + 3 * y) % 5, x
        rot = int(ks.ROT[a][bcol])
        return ap(bcol, a, (z + 64 - rot) % 64)
    for y in range(5):
        for x in range(5):
            for l in range(4):
                acc = b.const(0)
                for z in reversed(range(16 * l, 16 * l + 16)):
                    acc = acc + acc + _xor(bb(x, y, z), _andn(bb((x + 1) % 5, y, z), bb((x + 2) % 5, y, z)))
                b.assert_eq(acc, app(y, x, l))
    for l in range(4):
        acc = b.const(0)
        for z in reversed(range(16 * l, 16 * l + 16)):
            _assert_bool(b, bits00(z))
            acc = acc + acc + bits00(z)
        b.assert_eq(acc, app(0, 0, l))

    def xored_bit(i):
        rc_bit = b.const(0)
        for r in range(ROUNDS):
            rc_bit = rc_bit + flags[r] * ((int(ks.RC[r]) >> i) & 1)
        return _xor(bits00(i), rc_bit)
    for l in range(4):
        acc = b.const(0)
        for z in reversed(range(16 * l, 16 * l + 16)):
            acc = acc + acc + xored_bit(z)
        b.assert_eq(acc, appp00(l))
    for x in range(5):
        for y in range(5):
            for l in range(4):
                out = appp00(l) if (x == 0 and y == 0) else app(y, x, l)
                b.when_transition().when(not_final_step).assert_eq(out, nxt(a_col(y, x, l)))


def _eval_memory_access(b, shard, clk, addr, base, do_check, is_write):
    """eval_memory_access on the MemoryReadCols / MemoryWriteCols starting at column `base`."""
    m = b.main
    acc = base + (4 if is_write else 0)                    # MemoryWriteCols: prev_value word first
    value = [m(acc + C.M_VALUE + k) for k in range(4)]
    prev_value = [m(base + k) for k in range(4)] if is_write else value
    prev_shard, prev_clk, compare_clk = m(acc + C.M_PREV_SHARD), m(acc + C.M_PREV_CLK), m(acc + C.M_COMPARE_CLK)
    l16, l8 = m(acc + C.M_DIFF16), m(acc + C.M_DIFF8)
    _assert_bool(b, do_check)
    # eval_memory_access_timestamp
    b.when(do_check).assert_zero(compare_clk * (compare_clk - 1))
    b.when(do_check).when(compare_clk).assert_eq(shard, prev_shard)
    prev_comp = compare_clk * prev_clk + (1 - compare_clk) * prev_shard
    cur_comp = compare_clk * clk + (1 - compare_clk) * shard
    diff_minus_one = cur_comp - prev_comp - 1
    # eval_range_check_24bits
    b.when(do_check).assert_eq(diff_minus_one, l16 + l8 * (1 << 16))
    b.send(KIND_BYTE, [BYTE_U16RANGE, l16, 0, 0, 0], do_check)
    b.send(KIND_BYTE, [BYTE_U8RANGE, 0, 0, 0, l8], do_check)
    b.send(KIND_MEMORY, [prev_shard, prev_clk, addr] + prev_value, do_check)
    b.receive(KIND_MEMORY, [shard, clk, addr] + value, do_check)
    return value, prev_value


def _eval_keccak_sponge(b):
    m = b.main
    nxt = lambda c: b.main(c, next=True)
    first_block, final_block = m(C.IS_FIRST_INPUT_BLOCK), m(C.IS_FINAL_INPUT_BLOCK)
    first_step, final_step = m(C.STEP_FLAGS), m(C.STEP_FLAGS + ROUNDS - 1)
    not_final_step = 1 - final_step
    write_output, receive_syscall, read_block = m(C.WRITE_OUTPUT), m(C.RECEIVE_SYSCALL), m(C.READ_BLOCK)
    is_real, is_absorbed = m(C.IS_REAL), m(C.IS_ABSORBED)
    not_final_sponge = 1 - write_output
    shard, clk = m(C.SHARD), m(C.CLK)
    in_addr, out_addr, in_len, absorbed = m(C.INPUT_ADDRESS), m(C.OUTPUT_ADDRESS), m(C.INPUT_LEN), m(C.ALREADY_ABSORBED_U32S)
    # eval_flags
    b.assert_eq(first_block * first_step * is_real, receive_syscall)
    b.assert_eq(final_block * final_step * is_real, write_output)
    b.assert_eq(is_absorbed, final_step * (1 - final_block) * is_real)
    # eval_memory_access
    _eval_memory_access(b, shard, clk, out_addr + 64, C.INPUT_LENGTH_MEM, receive_syscall, False)
    for k in range(4):
        v = m(C.INPUT_LENGTH_MEM + k)
        b.when(is_real).assert_eq(v, v)
    for i in range(RATE):
        _eval_memory_access(b, shard, clk, in_addr + 4 * i, C.BLOCK_MEM + 9 * i, read_block, False)
    for i in range(RATE):
        for k in range(4):
            v = m(C.BLOCK_MEM + 9 * i + k)
            b.when(is_real).assert_eq(v, v)
    out_values = []
    for i in range(OUT):
        v, _ = _eval_memory_access(b, shard, clk + 1, out_addr + 4 * i, C.OUTPUT_MEM + 13 * i, write_output, True)
        out_values.append(v)
    # eval_state_keccakf
    word = lambda base, j, nx=False: [(nxt if nx else m)(base + 4 * j + k) for k in range(4)]

    def limbs_of(lo, hi):
        return [lo[0] + lo[1] * 256, lo[2] + lo[3] * 256, hi[0] + hi[1] * 256, hi[2] + hi[3] * 256]

    def appp(y, x, l):
        return m(C.A_PRIME_PRIME_PRIME_0_0_LIMBS + l) if (x == 0 and y == 0) else m(C.A_PRIME_PRIME + (y * 5 + x) * 4 + l)
    for i in range(STATE // 2):
        y, x = i // 5, i % 5
        src = C.XORED_GENERAL_RATE if i < RATE // 2 else C.ORIGINAL_STATE
        mem = limbs_of(word(src, 2 * i), word(src, 2 * i + 1))
        for j in range(4):
            b.when(first_step * is_real).assert_eq(mem[j], m(C.A + (y * 5 + x) * 4 + j))
        mem = limbs_of(word(C.ORIGINAL_STATE, 2 * i, True), word(C.ORIGINAL_STATE, 2 * i + 1, True))
        for j in range(4):
            b.when(is_absorbed).assert_eq(mem[j], appp(y, x, j))
    for i in range(OUT // 2):
        y, x = i // 5, i % 5
        mem = limbs_of(out_values[2 * i], out_values[2 * i + 1])
        for j in range(4):
            b.when(write_output).assert_eq(mem[j], appp(y, x, j))
    # syscall
    b.receive(KIND_SYSCALL, [shard, clk, SYSCALL_ID, in_addr, out_addr], receive_syscall)
    # the inputs stay the same throughout the rows of a sponge
    t = b.when_transition().when(not_final_sponge)
    for c in (C.SHARD, C.CLK, C.IS_REAL, C.INPUT_LEN, C.OUTPUT_ADDRESS):
        t.assert_eq(m(c), nxt(c))
    b.when_last_row().assert_zero(is_real)
    # XorOperation::eval
    for i in range(RATE):
        a, bw, val = word(C.ORIGINAL_STATE, i), [m(C.BLOCK_MEM + 9 * i + k) for k in range(4)], word(C.XORED_GENERAL_RATE, i)
        for k in range(4):
            b.send(KIND_BYTE, [BYTE_XOR, val[k], 0, a[k], bw[k]], read_block)
    # absorbed words
    b.when_transition().when(not_final_step).assert_eq(absorbed, nxt(C.ALREADY_ABSORBED_U32S))
    b.when(first_block).assert_eq(absorbed, 0)
    b.when(final_block).assert_eq(absorbed, in_len - RATE)
    b.when(is_absorbed).assert_eq(absorbed, nxt(C.ALREADY_ABSORBED_U32S) - RATE)
    b.when(is_absorbed).assert_eq(in_addr, nxt(C.INPUT_ADDRESS) - RATE * 4)
    _eval_p3_keccak(b)


def keccak_sponge_chip() -> Chip:
    return Chip("KeccakSponge", 0, C.WIDTH, _eval_keccak_sponge)


# ---- receivers --------------------------------------------------------------------------------------------------
BYTE_PREP_WIDTH, BYTE_MAIN_WIDTH = 4, 4


def byte_chip() -> Chip:
    """Byte table of 2^16 rows (b = row >> 8, c = row & 255).  Preprocessed: value_u16 (= the row index), b, c, xor;
    multiplicities: the one-value range lookups of the synthetic tables (ziren_b200/synthetic.py), XOR, U8Range, U16Range."""
    def ev(b):
        v16, bb, cc, xr = (b.prep(i) for i in range(4))
        b.receive(KIND_BYTE, [v16], b.main(0))
        b.receive(KIND_BYTE, [BYTE_XOR, xr, 0, bb, cc], b.main(1))
        b.receive(KIND_BYTE, [BYTE_U8RANGE, 0, 0, bb, cc], b.main(2))
        b.receive(KIND_BYTE, [BYTE_U16RANGE, v16, 0, 0, 0], b.main(3))
    return Chip("Byte", BYTE_PREP_WIDTH, BYTE_MAIN_WIDTH, ev, local_only=True)


def byte_prep() -> np.ndarray:
    idx = np.arange(1 << 16, dtype=np.uint32)
    return np.stack([idx, idx >> 8, idx & 255, (idx >> 8) ^ (idx & 255)], axis=1).astype(np.uint32)


def memory_local_chip() -> Chip:
    """One memory access per row: receives the tuple the access found (initial), sends the one it left (final)."""
    def ev(b):
        m = b.main
        is_real = m(13)
        _assert_bool(b, is_real)
        b.receive(KIND_MEMORY, [m(1), m(2), m(0)] + [m(3 + k) for k in range(4)], is_real)
        b.send(KIND_MEMORY, [m(7), m(8), m(0)] + [m(9 + k) for k in range(4)], is_real)
    return Chip("MemoryLocalPrecompile", 0, 14, ev)


def syscall_precompile_chip() -> Chip:
    def ev(b):
        m = b.main
        _assert_bool(b, m(5))
        b.send(KIND_SYSCALL, [m(0), m(1), m(2), m(3), m(4)], m(5))
    return Chip("SyscallPrecompile", 0, 6, ev)


def _bytes_of(words: np.ndarray) -> np.ndarray:
    return np.stack([(words >> np.uint32(8 * k)) & np.uint32(255) for k in range(4)], axis=-1)


def receiver_tables(blocks: np.ndarray):
    """What the rest of the machine holds for the KeccakSponge rows of `blocks` ((n, 384) records): the Byte
    multiplicities (XOR, U8Range, U16Range: `generate_dependencies`, trace.rs:30-56), one MemoryLocalPrecompile row per
    memory access and one SyscallPrecompile row per event.  Returns (byte_mults (65536, 3) uint64, memory rows, syscall rows)."""
    rec = np.ascontiguousarray(blocks, dtype=np.uint32).reshape(-1, ks.REC_WORDS)
    n = len(rec)
    mult = np.zeros((1 << 16, 3), np.uint64)
    if n == 0:
        return mult, np.zeros((0, 14), np.uint32), np.zeros((0, 6), np.uint32)
    first, last = rec[:, ks.R_BLOCK] == 0, rec[:, ks.R_BLOCK] + 1 == rec[:, ks.R_NBLOCKS]
    xored = rec[:, ks.R_XORED_STATE:ks.R_XORED_STATE + RATE]
    inp = rec[:, ks.R_INPUT:ks.R_INPUT + RATE]
    orig = xored ^ inp
    idx = (_bytes_of(orig).astype(np.int64) << 8) | _bytes_of(inp).astype(np.int64)
    mult[:, 0] = np.bincount(idx.ravel(), minlength=1 << 16)
    reads = rec[:, ks.R_READS:ks.R_READS + 5 * RATE].reshape(n, RATE, 5)
    addr_r = rec[:, ks.R_INPUT_ADDR][:, None] + (rec[:, ks.R_BLOCK][:, None] * RATE + np.arange(RATE, dtype=np.uint32)[None, :]) * 4
    len_r = rec[first, ks.R_LEN_READ:ks.R_LEN_READ + 5]
    writes = rec[last, ks.R_WRITES:ks.R_WRITES + 6 * OUT].reshape(-1, OUT, 6)
    addr_w = rec[last, ks.R_OUTPUT_ADDR][:, None] + np.arange(OUT, dtype=np.uint32)[None, :] * 4
    # (addr, value, shard, ts, prev_value, prev_shard, prev_ts) of every access
    acc = np.concatenate([
        np.stack([addr_r.ravel(), reads[:, :, 0].ravel(), reads[:, :, 1].ravel(), reads[:, :, 2].ravel(), reads[:, :, 0].ravel(),
                  reads[:, :, 3].ravel(), reads[:, :, 4].ravel()], axis=1),
        np.stack([rec[first, ks.R_OUTPUT_ADDR] + 64, len_r[:, 0], len_r[:, 1], len_r[:, 2], len_r[:, 0], len_r[:, 3], len_r[:, 4]], axis=1),
        np.stack([addr_w.ravel(), writes[:, :, 0].ravel(), writes[:, :, 1].ravel(), writes[:, :, 2].ravel(), writes[:, :, 3].ravel(),
                  writes[:, :, 4].ravel(), writes[:, :, 5].ravel()], axis=1)]).astype(np.uint32)
    same = acc[:, 5] == acc[:, 2]
    d = (np.where(same, acc[:, 3] - acc[:, 6], acc[:, 2] - acc[:, 5]) - np.uint32(1)).astype(np.uint32)
    mult[:, 2] = np.bincount((d & np.uint32(0xFFFF)).astype(np.int64), minlength=1 << 16)
    mult[:, 1] = np.bincount(((d >> np.uint32(16)) & np.uint32(0xFF)).astype(np.int64), minlength=1 << 16)
    mem = np.zeros((len(acc), 14), np.uint32)
    mem[:, 0] = acc[:, 0]
    mem[:, 1], mem[:, 2] = acc[:, 5], acc[:, 6]
    mem[:, 3:7] = _bytes_of(acc[:, 4])
    mem[:, 7], mem[:, 8] = acc[:, 2], acc[:, 3]
    mem[:, 9:13] = _bytes_of(acc[:, 1])
    mem[:, 13] = 1
    ev = rec[first]
    sysc = np.stack([ev[:, ks.R_SHARD], ev[:, ks.R_CLK], np.full(len(ev), SYSCALL_ID, np.uint32), ev[:, ks.R_INPUT_ADDR],
                     ev[:, ks.R_OUTPUT_ADDR], np.ones(len(ev), np.uint32)], axis=1).astype(np.uint32)
    return mult, mem, sysc


def pad_rows(rows: np.ndarray, min_log: int = 4) -> np.ndarray:
    n = max(len(rows), 1)
    h = 1 << max(min_log, int(n - 1).bit_length())
    out = np.zeros((h, rows.shape[1]), np.uint32)
    out[:len(rows)] = rows
    return out

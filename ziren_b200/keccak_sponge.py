"""Host side of GPU trace generation for the KeccakSponge precompile chip:
the column map of `KeccakSpongeCols`, the flat per-block event record `zkb200_generate_keccak_sponge_trace`
takes (include/zkb200.h, `zkb200_keccak_block`), and well-formed synthetic events (no guest ELF can be
executed here).

Reference: columns crates/core/machine/src/syscall/precompiles/keccak_sponge/columns.rs:14-37 with
`p3_keccak_air::KeccakCols` in front; rows trace.rs:58-196; event
crates/core/executor/src/events/precompiles/keccak_sponge.rs:15-46; memory records
crates/core/executor/src/events/memory.rs:46-82.  The Rust event holds Vecs, so the shim flattens it into
one fixed-size record per absorbed block (24 rows each)."""
from __future__ import annotations

import numpy as np

RATE_U32S, STATE_U32S, OUTPUT_U32S, NUM_ROUNDS = 36, 50, 16, 24
REC_WORDS = 384
# record layout (32-bit words)
R_SHARD, R_CLK, R_INPUT_ADDR, R_OUTPUT_ADDR, R_INPUT_LEN, R_BLOCK, R_NBLOCKS = range(7)
R_XORED_STATE, R_INPUT, R_READS, R_LEN_READ, R_WRITES = 8, 58, 94, 274, 279


class Cols:
    """Column offsets of KeccakSpongeCols (KeccakCols first)."""
    STEP_FLAGS, EXPORT, PREIMAGE, A, C, C_PRIME, A_PRIME, A_PRIME_PRIME = 0, 24, 25, 125, 225, 545, 865, 2465
    A_PRIME_PRIME_0_0_BITS, A_PRIME_PRIME_PRIME_0_0_LIMBS, NUM_KECCAK_COLS = 2565, 2629, 2633
    BLOCK_MEM = 2633                       # 36 x MemoryReadCols (9)
    SHARD, CLK, IS_REAL, READ_BLOCK, INPUT_ADDRESS, OUTPUT_ADDRESS, INPUT_LEN, ALREADY_ABSORBED_U32S = range(2957, 2965)
    IS_ABSORBED, RECEIVE_SYSCALL, WRITE_OUTPUT, IS_FIRST_INPUT_BLOCK, IS_FINAL_INPUT_BLOCK = range(2965, 2970)
    ORIGINAL_STATE = 2970                  # 50 x Word
    XORED_GENERAL_RATE = 3170              # 36 x XorOperation (Word)
    INPUT_LENGTH_MEM = 3314                # MemoryReadCols
    OUTPUT_MEM = 3323                      # 16 x MemoryWriteCols (13)
    WIDTH = 3531
    # MemoryAccessCols: value[4], prev_shard, prev_clk, compare_clk, diff_16bit_limb, diff_8bit_limb
    M_VALUE, M_PREV_SHARD, M_PREV_CLK, M_COMPARE_CLK, M_DIFF16, M_DIFF8 = 0, 4, 5, 6, 7, 8


WIDTH = Cols.WIDTH

# keccak-air/src/constants.rs
ROT = np.array([[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]])  # [x][y]
RC = np.array([0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
               0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
               0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
               0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
               0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008], dtype=np.uint64)


def _rotl(v, r):
    r = int(r)
    return v if r == 0 else (v << np.uint64(r)) | (v >> np.uint64(64 - r))


def keccak_f1600(states: np.ndarray) -> np.ndarray:
    """Keccak-f[1600] on (n, 25) uint64 states, lane (x, y) at index x + 5y."""
    a = [states[:, i].copy() for i in range(25)]
    for rnd in range(24):
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x + 4) % 5] ^ _rotl(c[(x + 1) % 5], 1) for x in range(5)]
        a = [a[i] ^ d[i % 5] for i in range(25)]
        b = [None] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rotl(a[x + 5 * y], ROT[x][y])
        a = [b[i] ^ (~b[(i % 5 + 1) % 5 + 5 * (i // 5)] & b[(i % 5 + 2) % 5 + 5 * (i // 5)]) for i in range(25)]
        a[0] = a[0] ^ RC[rnd]
    return np.stack(a, axis=1)


def padded_log_height(n_blocks: int, fixed_log2_rows: int | None = None) -> int:
    """Rows are padded to `num_real_rows.next_power_of_two()` (trace.rs:86), or to the shape's fixed height."""
    rows = n_blocks * NUM_ROUNDS
    if fixed_log2_rows is not None:
        if rows > (1 << fixed_log2_rows):
            raise ValueError(f"fixed log2 rows is too small: got {rows} rows, expected at most {1 << fixed_log2_rows}")
        return fixed_log2_rows
    return int(rows - 1).bit_length() if rows > 1 else 0


def synthetic_blocks(n_events: int, blocks_per_event, seed: int = 0, shard: int = 1) -> np.ndarray:
    """Block records of `n_events` well-formed KeccakSpongeEvents (what the MIPS executor would have recorded):
    seeded random input words, the sponge state chained through Keccak-f between the blocks of an event, memory
    records whose previous access lies earlier in the same shard (clk comparison) or in an earlier shard.
    `blocks_per_event`: int or sequence of ints.  Returns (total_blocks, 384) uint32."""
    rng = np.random.default_rng(0x5EC0 + seed)
    per = np.broadcast_to(np.asarray(blocks_per_event, dtype=np.int64), (n_events,)).copy() if n_events else np.zeros(0, np.int64)
    total = int(per.sum())
    rec = np.zeros((total, REC_WORDS), np.uint32)
    if total == 0:
        return rec
    first = np.concatenate([[0], np.cumsum(per)[:-1]])
    ev_of = np.repeat(np.arange(n_events), per)
    blk = np.arange(total) - first[ev_of]
    # per event
    clk = (5 + 8 * np.arange(1, n_events + 1) * 3).astype(np.uint32)          # increasing, < 2^24 apart from prev accesses
    in_addr = (rng.integers(1 << 16, 1 << 28, n_events) & ~np.int64(3)).astype(np.uint32)
    out_addr = (rng.integers(1 << 28, 1 << 29, n_events) & ~np.int64(3)).astype(np.uint32)
    rec[:, R_SHARD] = shard
    rec[:, R_CLK] = clk[ev_of]
    rec[:, R_INPUT_ADDR] = in_addr[ev_of]
    rec[:, R_OUTPUT_ADDR] = out_addr[ev_of]
    rec[:, R_INPUT_LEN] = (per[ev_of] * RATE_U32S).astype(np.uint32)
    rec[:, R_BLOCK] = blk.astype(np.uint32)
    rec[:, R_NBLOCKS] = per[ev_of].astype(np.uint32)
    inp = rng.integers(0, 1 << 32, (total, RATE_U32S), dtype=np.uint64).astype(np.uint32)
    rec[:, R_INPUT:R_INPUT + RATE_U32S] = inp
    # sponge chain, all events in lock step over the block index
    state = np.zeros((n_events, STATE_U32S), np.uint32)
    final_state = np.zeros((n_events, STATE_U32S), np.uint32)
    for b in range(int(per.max())):
        live = np.nonzero(per > b)[0]
        rows = first[live] + b
        x = state[live].copy()
        x[:, :RATE_U32S] ^= inp[rows]
        rec[rows, R_XORED_STATE:R_XORED_STATE + STATE_U32S] = x
        x64 = x[:, 0::2].astype(np.uint64) | (x[:, 1::2].astype(np.uint64) << np.uint64(32))
        y64 = keccak_f1600(x64)
        y = np.empty_like(x)
        y[:, 0::2] = (y64 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        y[:, 1::2] = (y64 >> np.uint64(32)).astype(np.uint32)
        state[live] = y
        done = live[per[live] == b + 1]
        final_state[done] = state[done]
    # memory records: previous access one to 2^20 cycles earlier in this shard, or (one in eight) in an earlier shard
    def prev_of(n, cur_clk):
        earlier_shard = (rng.integers(0, 8, n) == 0) & (shard > 1)
        pshard = np.where(earlier_shard, rng.integers(1, max(shard, 2), n), shard).astype(np.uint32)
        back = rng.integers(1, np.minimum(cur_clk.astype(np.int64), 1 << 20) + 1, n)
        pclk = np.where(earlier_shard, rng.integers(0, 1 << 22, n), cur_clk.astype(np.int64) - back).astype(np.uint32)
        return pshard, pclk
    reads = rec[:, R_READS:R_READS + 5 * RATE_U32S].reshape(total, RATE_U32S, 5)
    ck = np.repeat(rec[:, R_CLK], RATE_U32S)
    ps, pc = prev_of(total * RATE_U32S, ck)
    reads[:, :, 0] = inp
    reads[:, :, 1] = shard
    reads[:, :, 2] = rec[:, R_CLK][:, None]
    reads[:, :, 3] = ps.reshape(total, RATE_U32S)
    reads[:, :, 4] = pc.reshape(total, RATE_U32S)
    rec[:, R_READS:R_READS + 5 * RATE_U32S] = reads.reshape(total, -1)
    ps, pc = prev_of(total, rec[:, R_CLK])
    rec[:, R_LEN_READ + 0] = rec[:, R_INPUT_LEN]
    rec[:, R_LEN_READ + 1] = shard
    rec[:, R_LEN_READ + 2] = rec[:, R_CLK]
    rec[:, R_LEN_READ + 3] = ps
    rec[:, R_LEN_READ + 4] = pc
    writes = np.zeros((total, OUTPUT_U32S, 6), np.uint32)
    ps, pc = prev_of(total * OUTPUT_U32S, np.repeat(rec[:, R_CLK] + 1, OUTPUT_U32S))
    writes[:, :, 0] = final_state[ev_of][:, :OUTPUT_U32S]
    writes[:, :, 1] = shard
    writes[:, :, 2] = (rec[:, R_CLK] + 1)[:, None]
    writes[:, :, 3] = rng.integers(0, 1 << 32, (total, OUTPUT_U32S), dtype=np.uint64).astype(np.uint32)
    writes[:, :, 4] = ps.reshape(total, OUTPUT_U32S)
    writes[:, :, 5] = pc.reshape(total, OUTPUT_U32S)
    rec[:, R_WRITES:R_WRITES + 6 * OUTPUT_U32S] = writes.reshape(total, -1)
    return rec

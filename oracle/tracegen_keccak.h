// CPU restatement of the KeccakSponge chip's trace generation — TEST
// INFRASTRUCTURE, the checker for ziren_b200/csrc/tracegen_keccak.cuh, never the thing shipped or measured.
//
// What it follows:
//   * KeccakSpongeChip::{generate_trace, event_to_rows}, crates/core/machine/src/syscall/precompiles/
//     keccak_sponge/trace.rs:58-196; column order of KeccakSpongeCols, keccak_sponge/columns.rs:14-37;
//     MemoryReadCols / MemoryWriteCols / MemoryAccessCols, crates/core/machine/src/memory/consistency/
//     columns.rs:4-51 and their populate, memory/consistency/trace.rs:8-103 (C++ twin:
//     crates/core/machine/include/memory.hpp:10-52, compiled into oracle/_ref/libzkref_core.so and compared
//     with ks_mem_access below); XorOperation::populate, crates/core/machine/src/operations/xor.rs:18-38.
//   * the 2633 permutation columns are p3_keccak_air::generate_trace_rows (trace.rs:13,134), which lives in
//     the UN-VENDORED dependency ProjectZKM/Plonky3 @ faa24ca4597eebeecbf71b194b71c7d1a99b3f01
//     (keccak-air/src/{columns,generation,constants}.rs; Cargo.lock).  Its published algorithm is restated
//     here the way that crate writes it: every round works on the 16-bit limb and bit COLUMNS of the row
//     (C = column parity bits, C' = C ^ C[x-1] ^ rot(C[x+1]), A' = A ^ C ^ C', B = the aliased rotation of
//     A', A'' = B ^ (~B[x+1] & B[x+2]) packed to limbs, A''' = A''[0][0] ^ RC), a row's `a` being the
//     previous row's a_prime_prime_prime.  `export` is never set by generate_trace_rows.
// PARITY: the memory columns are pinned by the reference's own C++; the permutation columns are pinned by
// known answers only (SHA3-256 of hashlib through these rows, tests/test_tracegen_keccak.py) and by the
// restated AIR accepting the trace — there is no reference source or golden row for them on disk.
//
// Event encoding (ours: the Rust event holds Vecs): one record of KS_REC_WORDS words per absorbed BLOCK
// (= 24 rows), see include/zkb200.h `zkb200_keccak_block`.
#pragma once
#include <cstring>
#include <vector>
#include "kb.h"

namespace zko {

enum { KS_RATE = 36, KS_STATE = 50, KS_OUT = 16, KS_ROUNDS = 24, KS_REC_WORDS = 384 };
enum { KS_FLAGS = 0, KS_EXPORT = 24, KS_PRE = 25, KS_A = 125, KS_C = 225, KS_CP = 545, KS_AP = 865, KS_APP = 2465,
       KS_APP00_BITS = 2565, KS_APPP00 = 2629, KS_P3_COLS = 2633,
       KS_BLOCK_MEM = 2633, KS_SHARD = 2957, KS_CLK, KS_IS_REAL, KS_READ_BLOCK, KS_INPUT_ADDR, KS_OUTPUT_ADDR, KS_INPUT_LEN,
       KS_ABSORBED_U32S, KS_IS_ABSORBED, KS_RECEIVE_SYSCALL, KS_WRITE_OUTPUT, KS_IS_FIRST, KS_IS_FINAL,
       KS_ORIG_STATE = 2970, KS_XORED = 3170, KS_LEN_MEM = 3314, KS_OUT_MEM = 3323, KS_WIDTH = 3531 };
// record layout (words)
enum { KR_SHARD = 0, KR_CLK, KR_INPUT_ADDR, KR_OUTPUT_ADDR, KR_INPUT_LEN, KR_BLOCK, KR_NBLOCKS, KR_RESERVED,
       KR_XORED_STATE = 8, KR_INPUT = 58, KR_READS = 94, KR_LEN_READ = 274, KR_WRITES = 279 };

// keccak-air/src/constants.rs: rotation offsets R[x][y] and the 24 round constants
static const unsigned char KS_R[5][5] = {{0, 36, 3, 41, 18}, {1, 44, 10, 45, 2}, {62, 6, 43, 15, 61}, {28, 55, 25, 21, 56}, {27, 20, 39, 8, 14}};
static const unsigned long long KS_RC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
    0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};

// one round on the row's own columns (generation.rs generate_trace_row_for_round)
static inline void ks_p3_round(u32* r, int round) {
  r[KS_FLAGS + round] = 1;
  auto a_bit = [&](int y, int x, int z) { return (r[KS_A + (y * 5 + x) * 4 + z / 16] >> (z % 16)) & 1u; };
  for (int x = 0; x < 5; x++)
    for (int z = 0; z < 64; z++) {
      u32 v = 0;
      for (int y = 0; y < 5; y++) v ^= a_bit(y, x, z);
      r[KS_C + x * 64 + z] = v;
    }
  for (int x = 0; x < 5; x++)
    for (int z = 0; z < 64; z++)
      r[KS_CP + x * 64 + z] = r[KS_C + x * 64 + z] ^ r[KS_C + ((x + 4) % 5) * 64 + z] ^ r[KS_C + ((x + 1) % 5) * 64 + (z + 63) % 64];
  for (int y = 0; y < 5; y++)
    for (int x = 0; x < 5; x++)
      for (int z = 0; z < 64; z++)
        r[KS_AP + (y * 5 + x) * 64 + z] = a_bit(y, x, z) ^ r[KS_C + x * 64 + z] ^ r[KS_CP + x * 64 + z];
  // columns.rs b(): B[x, y] = ROT(A'[(x + 3y) % 5, x], R[(x + 3y) % 5][x])
  auto b = [&](int x, int y, int z) {
    const int a = (x + 3 * y) % 5, bb = x, rot = KS_R[a][bb];
    return r[KS_AP + (bb * 5 + a) * 64 + (z + 64 - rot) % 64];
  };
  for (int y = 0; y < 5; y++)
    for (int x = 0; x < 5; x++)
      for (int limb = 0; limb < 4; limb++) {
        u32 v = 0;
        for (int k = 0; k < 16; k++) {
          const int z = limb * 16 + k;
          const u32 bit = b(x, y, z) ^ ((1u ^ b((x + 1) % 5, y, z)) & b((x + 2) % 5, y, z));
          v |= bit << k;
        }
        r[KS_APP + (y * 5 + x) * 4 + limb] = v;
      }
  for (int z = 0; z < 64; z++) r[KS_APP00_BITS + z] = (r[KS_APP + z / 16] >> (z % 16)) & 1u;
  for (int limb = 0; limb < 4; limb++)
    r[KS_APPP00 + limb] = r[KS_APP + limb] ^ (u32)((KS_RC[round] >> (16 * limb)) & 0xffffu);
}
static inline u32 ks_appp(const u32* r, int y, int x, int limb) {
  return (y == 0 && x == 0) ? r[KS_APPP00 + limb] : r[KS_APP + (y * 5 + x) * 4 + limb];
}
// generate_trace_rows_for_perm: 24 rows of `pitch` words; only the first 2633 columns are written
static inline void ks_p3_rows(const unsigned long long input[25], u32* rows, size_t pitch) {
  for (int round = 0; round < KS_ROUNDS; round++) {
    u32* r = rows + (size_t)round * pitch;
    memset(r, 0, KS_P3_COLS * sizeof(u32));
    for (int i = 0; i < 25; i++)
      for (int limb = 0; limb < 4; limb++) r[KS_PRE + i * 4 + limb] = (u32)((input[i] >> (16 * limb)) & 0xffffu);
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++)
        for (int limb = 0; limb < 4; limb++)
          r[KS_A + (y * 5 + x) * 4 + limb] = round == 0 ? r[KS_PRE + (y * 5 + x) * 4 + limb] : ks_appp(r - pitch, y, x, limb);
    ks_p3_round(r, round);
  }
}

// MemoryAccessCols::populate_access (memory/consistency/trace.rs:73-103): value[4], prev_shard, prev_clk,
// compare_clk, diff_16bit_limb, diff_8bit_limb
static inline void ks_mem_access(u32* c, u32 value, u32 shard, u32 ts, u32 prev_shard, u32 prev_ts) {
  for (int i = 0; i < 4; i++) c[i] = (value >> (8 * i)) & 0xffu;
  c[4] = prev_shard % P;
  c[5] = prev_ts % P;
  const bool same = prev_shard == shard;
  c[6] = same ? 1 : 0;
  const u32 diff_minus_one = (same ? ts - prev_ts : shard - prev_shard) - 1u;
  c[7] = diff_minus_one & 0xffffu;
  c[8] = (diff_minus_one >> 16) & 0xffu;
}

// the 24 rows of one block record (event_to_rows, one iteration of its block loop)
static inline void ks_block_rows(const u32* rec, u32* rows) {
  unsigned long long xored[25];
  for (int i = 0; i < 25; i++) xored[i] = (unsigned long long)rec[KR_XORED_STATE + 2 * i] | ((unsigned long long)rec[KR_XORED_STATE + 2 * i + 1] << 32);
  ks_p3_rows(xored, rows, KS_WIDTH);
  const u32 i = rec[KR_BLOCK], nb = rec[KR_NBLOCKS];
  for (int round = 0; round < KS_ROUNDS; round++) {
    u32* r = rows + (size_t)round * KS_WIDTH;
    memset(r + KS_P3_COLS, 0, (KS_WIDTH - KS_P3_COLS) * sizeof(u32));
    const bool first_round = round == 0, last_round = round == KS_ROUNDS - 1, first = i == 0, last = i == nb - 1;
    r[KS_SHARD] = rec[KR_SHARD] % P;
    r[KS_CLK] = rec[KR_CLK] % P;
    r[KS_IS_REAL] = 1;
    r[KS_INPUT_LEN] = rec[KR_INPUT_LEN] % P;
    r[KS_ABSORBED_U32S] = (i * KS_RATE) % P;
    r[KS_IS_ABSORBED] = last_round && !last;
    r[KS_IS_FIRST] = first;
    r[KS_IS_FINAL] = last;
    r[KS_READ_BLOCK] = first_round;
    r[KS_RECEIVE_SYSCALL] = first && first_round;
    r[KS_WRITE_OUTPUT] = last && last_round;
    r[KS_OUTPUT_ADDR] = rec[KR_OUTPUT_ADDR] % P;
    r[KS_INPUT_ADDR] = (rec[KR_INPUT_ADDR] + i * KS_RATE * 4) % P;
    if (first_round)
      for (int j = 0; j < KS_RATE; j++) {
        const u32* m = rec + KR_READS + 5 * j;      // MemoryReadRecord {value, shard, timestamp, prev_shard, prev_timestamp}
        ks_mem_access(r + KS_BLOCK_MEM + 9 * j, m[0], m[1], m[2], m[3], m[4]);
      }
    for (int j = 0; j < KS_STATE; j++) {
      // state_u32s before this block is absorbed: the xored state with the block's words taken out again
      const u32 st = j < KS_RATE ? rec[KR_XORED_STATE + j] ^ rec[KR_INPUT + j] : rec[KR_XORED_STATE + j];
      for (int k = 0; k < 4; k++) r[KS_ORIG_STATE + 4 * j + k] = (st >> (8 * k)) & 0xffu;
    }
    if (first_round)
      for (int j = 0; j < KS_RATE; j++)
        for (int k = 0; k < 4; k++) r[KS_XORED + 4 * j + k] = (rec[KR_XORED_STATE + j] >> (8 * k)) & 0xffu;
    if (first && first_round) {
      const u32* m = rec + KR_LEN_READ;
      ks_mem_access(r + KS_LEN_MEM, m[0], m[1], m[2], m[3], m[4]);
    }
    if (last && last_round)
      for (int j = 0; j < KS_OUT; j++) {
        const u32* m = rec + KR_WRITES + 6 * j;     // MemoryWriteRecord {value, shard, timestamp, prev_value, prev_shard, prev_timestamp}
        u32* c = r + KS_OUT_MEM + 13 * j;
        for (int k = 0; k < 4; k++) c[k] = (m[3] >> (8 * k)) & 0xffu;
        ks_mem_access(c + 4, m[0], m[1], m[2], m[4], m[5]);
      }
  }
}

// generate_trace: rows of the blocks in order, then dummy rows (the zero-input permutation's rows, by
// row index mod 24, every sponge column zero) up to `height`
static inline void keccak_sponge_trace(const u32* recs, size_t n_blocks, size_t height, u32* out) {
  if (n_blocks * KS_ROUNDS > height) throw std::runtime_error("oracle: keccak_sponge_trace: more rows than height");
#pragma omp parallel for schedule(dynamic, 8)
  for (long long b = 0; b < (long long)n_blocks; b++) ks_block_rows(recs + (size_t)b * KS_REC_WORDS, out + (size_t)b * KS_ROUNDS * KS_WIDTH);
  std::vector<u32> dummy((size_t)KS_ROUNDS * KS_WIDTH, 0);
  unsigned long long zero[25] = {0};
  ks_p3_rows(zero, dummy.data(), KS_WIDTH);
  for (size_t i = n_blocks * KS_ROUNDS; i < height; i++) memcpy(out + i * KS_WIDTH, dummy.data() + (i % KS_ROUNDS) * KS_WIDTH, KS_WIDTH * sizeof(u32));
}

}  // namespace zko

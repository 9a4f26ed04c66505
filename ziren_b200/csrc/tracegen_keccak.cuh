// Row filler of the KeccakSponge precompile chip (the table that carries 88 % of a
// keccak-heavy shard's trace bytes).  One absorbed block of a KeccakSpongeEvent is 24 rows, one per Keccak-f round;
// row (block b, round r) is a pure function of the block's record and r, so every row is filled independently:
// the thread re-runs rounds 0..r-1 on the 25 64-bit lanes in registers (at most 23 rounds of ~170 word operations
// against the 3531 stores the row costs) and then writes the round's columns from whole-word values - C, C', A'
// as the bits of 64-bit lanes, A'' as 16-bit limbs - where the reference (and the oracle) walk bit and limb
// columns.  What each column holds follows KeccakSpongeChip::event_to_rows
// (crates/core/machine/src/syscall/precompiles/keccak_sponge/trace.rs:101-196), the column order of
// KeccakSpongeCols (keccak_sponge/columns.rs:14-37) behind p3_keccak_air::KeccakCols (un-vendored Plonky3
// keccak-air: step_flags[24], export, preimage[5][5][4], a[5][5][4], c[5][64], c_prime[5][64],
// a_prime[5][5][64], a_prime_prime[5][5][4], a_prime_prime_0_0_bits[64], a_prime_prime_prime_0_0_limbs[4]),
// MemoryAccessCols::populate_access (memory/consistency/trace.rs:73-103; C++ twin include/memory.hpp:10-34) and
// XorOperation::populate (operations/xor.rs:18-38).  Host and device run the same code (tests/hostcheck).
//
// Event record (ours; the Rust event holds Vecs): KS_REC_WORDS 32-bit words per block, include/zkb200.h.
#pragma once
#include "kb31.cuh"

namespace zkb {

constexpr int KS_ROUNDS = 24, KS_RATE = 36, KS_STATE = 50, KS_OUT = 16, KS_REC_WORDS = 384, KS_WIDTH = 3531;
// column offsets
constexpr int KS_FLAGS = 0, KS_EXPORT = 24, KS_PRE = 25, KS_A = 125, KS_C = 225, KS_CP = 545, KS_AP = 865, KS_APP = 2465,
              KS_APP00_BITS = 2565, KS_APPP00 = 2629, KS_BLOCK_MEM = 2633, KS_SHARD = 2957, KS_CLK = 2958, KS_IS_REAL = 2959,
              KS_READ_BLOCK = 2960, KS_INPUT_ADDR = 2961, KS_OUTPUT_ADDR = 2962, KS_INPUT_LEN = 2963, KS_ABSORBED_U32S = 2964,
              KS_IS_ABSORBED = 2965, KS_RECEIVE_SYSCALL = 2966, KS_WRITE_OUTPUT = 2967, KS_IS_FIRST = 2968, KS_IS_FINAL = 2969,
              KS_ORIG_STATE = 2970, KS_XORED = 3170, KS_LEN_MEM = 3314, KS_OUT_MEM = 3323;
// record offsets (words)
constexpr int KR_SHARD = 0, KR_CLK = 1, KR_INPUT_ADDR = 2, KR_OUTPUT_ADDR = 3, KR_INPUT_LEN = 4, KR_BLOCK = 5, KR_NBLOCKS = 6,
              KR_XORED_STATE = 8, KR_INPUT = 58, KR_READS = 94, KR_LEN_READ = 274, KR_WRITES = 279;

// iota constants; rho offsets by lane index x + 5y (keccak-air constants.rs RC, R[x][y])
#define KS_RC_LIST                                                                                                      \
  0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,   \
  0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,   \
  0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,   \
  0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,   \
  0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull
#if defined(__CUDACC__)
static __constant__ u64 ks_rc_dev[24] = {KS_RC_LIST};
#endif
static const u64 ks_rc_host[24] = {KS_RC_LIST};
KB_HD u64 ks_rc(u32 round) {
#if defined(__CUDA_ARCH__)
  return ks_rc_dev[round];
#else
  return ks_rc_host[round];
#endif
}
KB_HD constexpr int ks_rho(int lane) {
  constexpr int r[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  return r[lane];
}
KB_HD u64 ks_rotl(u64 v, int r) { return r ? (v << r) | (v >> (64 - r)) : v; }

// theta's column parities C and C' = C ^ C[x-1] ^ rot(C[x+1], 1)
KB_HD void ks_parities(const u64* a, u64* c, u64* cp) {
#pragma unroll
  for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
  for (int x = 0; x < 5; x++) cp[x] = c[x] ^ c[(x + 4) % 5] ^ ks_rotl(c[(x + 1) % 5], 1);
}
// rho + pi + chi on A' (in place): B[X, Y] = rot(A'[(X + 3Y) % 5, X]), A''[X, Y] = B ^ (~B[X+1] & B[X+2])
KB_HD void ks_rho_pi_chi(u64* a) {
  u64 b[25];
#pragma unroll
  for (int Y = 0; Y < 5; Y++)
#pragma unroll
    for (int X = 0; X < 5; X++) {
      const int src = (X + 3 * Y) % 5 + 5 * X;
      b[X + 5 * Y] = ks_rotl(a[src], ks_rho(src));
    }
#pragma unroll
  for (int Y = 0; Y < 5; Y++)
#pragma unroll
    for (int X = 0; X < 5; X++) a[X + 5 * Y] = b[X + 5 * Y] ^ (~b[(X + 1) % 5 + 5 * Y] & b[(X + 2) % 5 + 5 * Y]);
}
KB_HD void ks_round(u64* a, u64 rc) {
  u64 c[5], cp[5];
  ks_parities(a, c, cp);
#pragma unroll
  for (int i = 0; i < 25; i++) a[i] ^= c[i % 5] ^ cp[i % 5];
  ks_rho_pi_chi(a);
  a[0] ^= rc;
}

KB_HD u32 ks_f(u32 x) { return fp_from_canonical(x).v; }        // F::from_canonical_u32, x < p
KB_HD u32 ks_bit(u64 v, int z) { return ((u32)(v >> z) & 1u) ? KB_ONE : 0u; }

// MemoryAccessCols: value[4], prev_shard, prev_clk, compare_clk, diff_16bit_limb, diff_8bit_limb
template <class St>
KB_HD void ks_mem_access(St& st, int col, u32 value, u32 shard, u32 ts, u32 prev_shard, u32 prev_ts) {
#pragma unroll
  for (int i = 0; i < 4; i++) st(col + i, ks_f((value >> (8 * i)) & 0xffu));
  st(col + 4, ks_f(prev_shard));
  st(col + 5, ks_f(prev_ts));
  const bool same = prev_shard == shard;
  st(col + 6, same ? KB_ONE : 0u);
  const u32 d = (same ? ts - prev_ts : shard - prev_shard) - 1u;
  st(col + 7, ks_f(d & 0xffffu));
  st(col + 8, ks_f((d >> 16) & 0xffu));
}
template <class St>
KB_HD void ks_zero(St& st, int col, int n) { for (int i = 0; i < n; i++) st(col + i, 0u); }

// One row.  rec: the block's record, or nullptr for a padding row (the zero-input permutation's row of this round,
// every sponge column zero: trace.rs:75-91).  st(col, montgomery_word) stores one cell.
template <class St>
KB_HD void ks_fill_row(const u32* __restrict__ rec, u32 round, St& st) {
  u64 a[25];
#pragma unroll
  for (int i = 0; i < 25; i++) a[i] = rec ? ((u64)rec[KR_XORED_STATE + 2 * i] | ((u64)rec[KR_XORED_STATE + 2 * i + 1] << 32)) : 0ull;
  // ---- p3 keccak columns ----
  for (int i = 0; i < KS_ROUNDS; i++) st(KS_FLAGS + i, (u32)i == round ? KB_ONE : 0u);
  st(KS_EXPORT, 0u);
#pragma unroll
  for (int i = 0; i < 25; i++)
#pragma unroll
    for (int l = 0; l < 4; l++) st(KS_PRE + 4 * i + l, ks_f((u32)(a[i] >> (16 * l)) & 0xffffu));
  for (u32 r = 0; r < round; r++) ks_round(a, ks_rc(r));
#pragma unroll
  for (int i = 0; i < 25; i++)
#pragma unroll
    for (int l = 0; l < 4; l++) st(KS_A + 4 * i + l, ks_f((u32)(a[i] >> (16 * l)) & 0xffffu));
  u64 c[5], cp[5];
  ks_parities(a, c, cp);
#pragma unroll
  for (int x = 0; x < 5; x++) {
#pragma unroll 8
    for (int z = 0; z < 64; z++) st(KS_C + 64 * x + z, ks_bit(c[x], z));
  }
#pragma unroll
  for (int x = 0; x < 5; x++) {
#pragma unroll 8
    for (int z = 0; z < 64; z++) st(KS_CP + 64 * x + z, ks_bit(cp[x], z));
  }
#pragma unroll
  for (int i = 0; i < 25; i++) {
    a[i] ^= c[i % 5] ^ cp[i % 5];
#pragma unroll 8
    for (int z = 0; z < 64; z++) st(KS_AP + 64 * i + z, ks_bit(a[i], z));
  }
  ks_rho_pi_chi(a);
#pragma unroll
  for (int i = 0; i < 25; i++)
#pragma unroll
    for (int l = 0; l < 4; l++) st(KS_APP + 4 * i + l, ks_f((u32)(a[i] >> (16 * l)) & 0xffffu));
#pragma unroll 8
  for (int z = 0; z < 64; z++) st(KS_APP00_BITS + z, ks_bit(a[0], z));
  {
    const u64 v = a[0] ^ ks_rc(round);
#pragma unroll
    for (int l = 0; l < 4; l++) st(KS_APPP00 + l, ks_f((u32)(v >> (16 * l)) & 0xffffu));
  }
  // ---- sponge columns ----
  if (!rec) { ks_zero(st, KS_BLOCK_MEM, KS_WIDTH - KS_BLOCK_MEM); return; }
  const u32 blk = rec[KR_BLOCK], nb = rec[KR_NBLOCKS];
  const bool first_round = round == 0, last_round = round == KS_ROUNDS - 1, first = blk == 0, last = blk + 1 == nb;
  if (first_round) {
    for (int j = 0; j < KS_RATE; j++) {
      const u32* m = rec + KR_READS + 5 * j;
      ks_mem_access(st, KS_BLOCK_MEM + 9 * j, m[0], m[1], m[2], m[3], m[4]);
    }
  } else ks_zero(st, KS_BLOCK_MEM, 9 * KS_RATE);
  st(KS_SHARD, ks_f(rec[KR_SHARD]));
  st(KS_CLK, ks_f(rec[KR_CLK]));
  st(KS_IS_REAL, KB_ONE);
  st(KS_READ_BLOCK, first_round ? KB_ONE : 0u);
  st(KS_INPUT_ADDR, ks_f(rec[KR_INPUT_ADDR] + blk * (KS_RATE * 4)));
  st(KS_OUTPUT_ADDR, ks_f(rec[KR_OUTPUT_ADDR]));
  st(KS_INPUT_LEN, ks_f(rec[KR_INPUT_LEN]));
  st(KS_ABSORBED_U32S, ks_f(blk * KS_RATE));
  st(KS_IS_ABSORBED, last_round && !last ? KB_ONE : 0u);
  st(KS_RECEIVE_SYSCALL, first && first_round ? KB_ONE : 0u);
  st(KS_WRITE_OUTPUT, last && last_round ? KB_ONE : 0u);
  st(KS_IS_FIRST, first ? KB_ONE : 0u);
  st(KS_IS_FINAL, last ? KB_ONE : 0u);
  for (int j = 0; j < KS_STATE; j++) {
    // the sponge state before this block: the xored state with the block's words taken out again
    const u32 v = j < KS_RATE ? rec[KR_XORED_STATE + j] ^ rec[KR_INPUT + j] : rec[KR_XORED_STATE + j];
#pragma unroll
    for (int k = 0; k < 4; k++) st(KS_ORIG_STATE + 4 * j + k, ks_f((v >> (8 * k)) & 0xffu));
  }
  if (first_round) {
    for (int j = 0; j < KS_RATE; j++) {
      const u32 v = rec[KR_XORED_STATE + j];
#pragma unroll
      for (int k = 0; k < 4; k++) st(KS_XORED + 4 * j + k, ks_f((v >> (8 * k)) & 0xffu));
    }
  } else ks_zero(st, KS_XORED, 4 * KS_RATE);
  if (first && first_round) {
    const u32* m = rec + KR_LEN_READ;
    ks_mem_access(st, KS_LEN_MEM, m[0], m[1], m[2], m[3], m[4]);
  } else ks_zero(st, KS_LEN_MEM, 9);
  if (last && last_round) {
    for (int j = 0; j < KS_OUT; j++) {
      const u32* m = rec + KR_WRITES + 6 * j;
#pragma unroll
      for (int k = 0; k < 4; k++) st(KS_OUT_MEM + 13 * j + k, ks_f((m[3] >> (8 * k)) & 0xffu));
      ks_mem_access(st, KS_OUT_MEM + 13 * j + 4, m[0], m[1], m[2], m[4], m[5]);
    }
  } else ks_zero(st, KS_OUT_MEM, 13 * KS_OUT);
}

}  // namespace zkb

// Poseidon2-KoalaBear width 16 (x^3 S-box, 8 external + 13 internal rounds) on Montgomery
// residues, for device kernels and for the host-side challenger.
// Round schedule per crates/primitives/src/lib.rs:1107-1123; linear layers per
// crates/core/machine/src/operations/poseidon2/air.rs:12-72 (M4 "light" MDS + column sums for
// external rounds, diag * s + sum(s) for internal rounds).
//
// Instruction-level design (profiles/README.md, "Poseidon2 v2"): on sm_100 integer work issues on
// two pipes of 16 lanes per scheduler, the FMA-heavy pipe (IMAD*) and the ALU pipe (IADD3, LOP3,
// SHF, VIADDMNMX); IMAD.WIDE / IMAD.HI occupy the heavy pipe for two slots.  ptxas balances the two
// pipes by turning plain adds into IMAD.IADD, but it counts the wide multiplies as one slot and
// overloads the heavy pipe (89 % busy at 69 % issue).  The permutation therefore pins every add to a
// pipe: "heavy" adds are a * one + b with `one` read from constant memory (an IMAD ptxas cannot
// turn back), "ALU" adds are the three-input a + b - p1 (an IADD3 it cannot turn into IMAD),
// p1/p2/one/neg1/zero being run-time copies of p, p, 1, -1, 0.  Host and device run the same
// expressions.
#pragma once
#include "kb31.cuh"

namespace zkb {

struct P2Consts {
  u32 ext[8][16];   // external round constants (rows 0..3 and 17..20 of RC_16_30), Montgomery
  u32 in[13];       // internal round constants (rows 4..16, column 0), Montgomery
  u32 p1, p2;       // p, twice (distinct addresses so that (x - p1) + p2 is not folded)
  u32 one, neg1, zero;
};

// RC_16_30 of crates/primitives/src/lib.rs:563-1105 (canonical u32 values, reduced mod p on use)
inline P2Consts p2_make_consts() {
  static const u32 rc[30][16] = {
#include "p2_rc.inc"
  };
  P2Consts c;
  auto m = [](u32 x) { return fp_from_canonical(x % KB_P).v; };
  for (int r = 0; r < 4; r++)
    for (int i = 0; i < 16; i++) { c.ext[r][i] = m(rc[r][i]); c.ext[4 + r][i] = m(rc[17 + r][i]); }
  for (int r = 0; r < 13; r++) c.in[r] = m(rc[4 + r][0]);
  c.p1 = c.p2 = KB_P;
  c.one = 1;
  c.neg1 = 0xffffffffu;
  c.zero = 0;
  return c;
}
// host-side table, built once (hash.cu); the device copy lives in __constant__ memory there
const P2Consts& p2_host_consts();

// how many of the 88 pipe-neutral adds of an external round (16 round-constant adds, 16 final adds,
// 12 column-sum adds, 44 M4 adds, in this order of preference) go to the heavy pipe;
// -1 leaves the choice to ptxas
#ifndef ZKB_P2_NH
#define ZKB_P2_NH -1
#endif
// internal rounds: 1 = all pipe-neutral adds on the heavy pipe, 0 = all on the ALU pipe, -1 = ptxas
#ifndef ZKB_P2_INT_HEAVY
#define ZKB_P2_INT_HEAVY -1
#endif

// a + b mod p, pinned to a pipe (mode 1 heavy, 0 ALU, -1 unpinned)
KB_HD u32 p2_add(u32 a, u32 b, const P2Consts& C, int mode) {
  if (mode > 0) { u32 s = a * C.one + b; u32 t = s - KB_P; return t < s ? t : s; }
  if (mode == 0) { u32 t = a + b - C.p1; u32 u = t + C.p2; return u < t ? u : t; }
  u32 s = a + b; u32 t = s - KB_P; return t < s ? t : s;
}
// a - b mod p
KB_HD u32 p2_sub(u32 a, u32 b, const P2Consts& C, int mode) {
  u32 s;
  if (mode > 0) s = b * C.neg1 + a;
  else if (mode == 0) s = a - b + C.zero;
  else s = a - b;
  u32 t = s + KB_P;
  return t < s ? t : s;
}
KB_HD int p2_ext_mode(int site) { return ZKB_P2_NH < 0 ? -1 : (site < ZKB_P2_NH ? 1 : 0); }

// M4 = [[2,3,1,1],[1,2,3,1],[1,1,2,3],[3,1,1,2]] on block b (sites 44 + 4*op + b)
KB_HD void p2_m4(u32& x0, u32& x1, u32& x2, u32& x3, const P2Consts& C, int b) {
#define M4ADD(op, a, c) p2_add(a, c, C, p2_ext_mode(44 + 4 * (op) + b))
  u32 t01 = M4ADD(0, x0, x1), t23 = M4ADD(1, x2, x3);
  u32 t0123 = M4ADD(2, t01, t23);
  u32 t01123 = M4ADD(3, t0123, x1), t01233 = M4ADD(4, t0123, x3);
  u32 d0 = M4ADD(5, x0, x0), d2 = M4ADD(6, x2, x2);
  u32 n3 = M4ADD(7, t01233, d0);
  u32 n1 = M4ADD(8, t01123, d2);
  x0 = M4ADD(9, t01123, t01);
  x2 = M4ADD(10, t01233, t23);
  x1 = n1;
  x3 = n3;
#undef M4ADD
}
KB_HD void p2_external_linear(u32* s, const P2Consts& C) {
#pragma unroll
  for (int b = 0; b < 4; b++) p2_m4(s[4 * b], s[4 * b + 1], s[4 * b + 2], s[4 * b + 3], C, b);
  u32 sums[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    u32 a = p2_add(s[k], s[4 + k], C, p2_ext_mode(32 + 3 * k));
    u32 c = p2_add(s[8 + k], s[12 + k], C, p2_ext_mode(32 + 3 * k + 1));
    sums[k] = p2_add(a, c, C, p2_ext_mode(32 + 3 * k + 2));
  }
#pragma unroll
  for (int j = 0; j < 16; j++) s[j] = p2_add(s[j], sums[j & 3], C, p2_ext_mode(16 + j));
}
// x^3 for x in [0, p): the square stays in [0, 2p) (no correction), which the second product tolerates
KB_HD u32 p2_cube(u32 x) { return mont_reduce(mul_wide(mont_reduce_lazy(mul_wide(x, x)), x)); }

// v * 2^-k mod p for 1 <= k <= 24, as a signed value in (-p, 2^(31-k)) (wrapped in u32):
// p = 127 * 2^24 + 1, so 2^-k = -127 * 2^(24-k) and v * 2^-k = (v >> k) - (v mod 2^k) * 127 * 2^(24-k)
KB_HD u32 p2_div2k_signed(u32 v, int k) {
  return (v >> k) - (v & ((1u << k) - 1)) * (127u << (24 - k));
}
// sum + v * 2^-k  /  sum - v * 2^-k, results in [0, p)
KB_HD u32 p2_add_div2k(u32 sum, u32 v, int k, const P2Consts& C, int mode) {
  u32 r = p2_div2k_signed(v, k);
  u32 r1 = r + KB_P; r = r1 < r ? r1 : r;                                            // [0, p)
  return p2_add(sum, r, C, mode);
}
KB_HD u32 p2_sub_div2k(u32 sum, u32 v, int k, const P2Consts& C, int mode) {
  u32 r = (v & ((1u << k) - 1)) * (127u << (24 - k)) + sum;                          // [0, 2p)
  u32 r1 = r - KB_P; r = r1 < r ? r1 : r;                                            // [0, p)
  u32 q = v >> k;
  u32 z = mode > 0 ? q * C.neg1 + r : (mode == 0 ? r - q + C.zero : r - q);          // (-2^(31-k), p)
  u32 z1 = z + KB_P; return z1 < z ? z1 : z;
}
// v / 2 in [0, p): (v >> 1) + (v & 1) * (p + 1) / 2
KB_HD u32 p2_halve(u32 v) { return (v & 1u) * ((KB_P + 1) >> 1) + (v >> 1); }

// s_i <- sum(s) + d_i * s_i with d = [-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 2^-8, 1/8, 2^-24, -2^-8, -1/8, -1/16, -2^-24]
KB_HD void p2_internal_linear(u32* s, const P2Consts& C) {
  const int M = ZKB_P2_INT_HEAVY;
#define IADD(a, b) p2_add(a, b, C, M)
#define ISUB(a, b) p2_sub(a, b, C, M)
  // s[0] (fresh out of the S-box) joins last: the rest of the tree overlaps the S-box latency
  u32 rest = IADD(IADD(IADD(IADD(s[1], s[2]), IADD(s[3], s[4])), IADD(IADD(s[5], s[6]), IADD(s[7], s[8]))),
                  IADD(IADD(IADD(s[9], s[10]), IADD(s[11], s[12])), IADD(IADD(s[13], s[14]), s[15])));
  u32 sum = IADD(rest, s[0]);
  s[0] = ISUB(ISUB(sum, s[0]), s[0]);
  s[1] = IADD(sum, s[1]);
  s[2] = IADD(IADD(sum, s[2]), s[2]);
  s[3] = IADD(sum, p2_halve(s[3]));
  { u32 d = IADD(s[4], s[4]); s[4] = IADD(IADD(sum, d), s[4]); }
  { u32 d = IADD(s[5], s[5]); s[5] = IADD(IADD(d, d), sum); }
  s[6] = ISUB(sum, p2_halve(s[6]));
  { u32 d = IADD(s[7], s[7]); s[7] = ISUB(ISUB(sum, d), s[7]); }
  { u32 d = IADD(s[8], s[8]); s[8] = ISUB(sum, IADD(d, d)); }
  s[9] = p2_add_div2k(sum, s[9], 8, C, M);
  s[10] = p2_add_div2k(sum, s[10], 3, C, M);
  s[11] = p2_add_div2k(sum, s[11], 24, C, M);
  s[12] = p2_sub_div2k(sum, s[12], 8, C, M);
  s[13] = p2_sub_div2k(sum, s[13], 3, C, M);
  s[14] = p2_sub_div2k(sum, s[14], 4, C, M);
  s[15] = p2_sub_div2k(sum, s[15], 24, C, M);
#undef IADD
#undef ISUB
}

// generic permutation over a constants table reachable from the calling side
KB_HD void p2_permute_with(Fp* st, const P2Consts& C) {
  u32 s[16];
#pragma unroll
  for (int i = 0; i < 16; i++) s[i] = st[i].v;
  p2_external_linear(s, C);
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = p2_cube(p2_add(s[i], C.ext[r][i], C, p2_ext_mode(i)));
    p2_external_linear(s, C);
  }
#pragma unroll 1
  for (int r = 0; r < 13; r++) {
    s[0] = p2_cube(p2_add(s[0], C.in[r], C, ZKB_P2_INT_HEAVY));
    p2_internal_linear(s, C);
  }
#pragma unroll 1
  for (int r = 4; r < 8; r++) {
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = p2_cube(p2_add(s[i], C.ext[r][i], C, p2_ext_mode(i)));
    p2_external_linear(s, C);
  }
#pragma unroll
  for (int i = 0; i < 16; i++) st[i].v = s[i];
}

inline void p2_permute_host(Fp* s) { p2_permute_with(s, p2_host_consts()); }

}  // namespace zkb

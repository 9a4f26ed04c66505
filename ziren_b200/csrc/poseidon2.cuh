// Poseidon2-KoalaBear width 16 (x^3 S-box, 8 external + 13 internal rounds) on Montgomery
// residues, for device kernels and for the host-side challenger.
// Round schedule per crates/primitives/src/lib.rs:1107-1123; linear layers per
// crates/core/machine/src/operations/poseidon2/air.rs:12-72 (M4 "light" MDS + column sums for
// external rounds, diag * s + sum(s) for internal rounds).
#pragma once
#include "kb31.cuh"

namespace zkb {

struct P2Consts {
  u32 ext[8][16];   // external round constants (rows 0..3 and 17..20 of RC_16_30), Montgomery
  u32 in[13];       // internal round constants (rows 4..16, column 0), Montgomery
  u32 diag[16];     // internal diagonal, Montgomery
  u32 big;          // 0xffffffff, see fp_add_alu
};

// host-side table, built once (hash.cu); the device copy lives in __constant__ memory there
const P2Consts& p2_host_consts();

#ifndef ZKB_STEER
#define ZKB_STEER 0   // measured: no gain at levels 1-3 (profiles/README.md), kept for the record
#endif
#if ZKB_STEER >= 1
#define P2_ADD1(a, b) fp_add_alu(a, b, big)
#else
#define P2_ADD1(a, b) ((a) + (b))
#endif
#if ZKB_STEER >= 2
#define P2_ADD2(a, b) fp_add_alu(a, b, big)
#else
#define P2_ADD2(a, b) ((a) + (b))
#endif
#if ZKB_STEER >= 3
#define P2_ADD3(a, b) fp_add_alu(a, b, big)
#else
#define P2_ADD3(a, b) ((a) + (b))
#endif
KB_HD void p2_m4(Fp& x0, Fp& x1, Fp& x2, Fp& x3, u32 big) {
  Fp t01 = P2_ADD1(x0, x1), t23 = P2_ADD1(x2, x3);
  Fp t0123 = P2_ADD1(t01, t23);
  Fp t01123 = P2_ADD2(t0123, x1), t01233 = P2_ADD2(t0123, x3);
  Fp n3 = t01233 + fp_double(x0);
  Fp n1 = t01123 + fp_double(x2);
  x0 = t01123 + t01;
  x2 = t01233 + t23;
  x1 = n1;
  x3 = n3;
}
KB_HD void p2_external_linear(Fp* s, u32 big) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) p2_m4(s[j], s[j + 1], s[j + 2], s[j + 3], big);
  Fp sums[4];
#pragma unroll
  for (int k = 0; k < 4; k++) sums[k] = P2_ADD2(P2_ADD2(s[k], s[4 + k]), P2_ADD2(s[8 + k], s[12 + k]));
#pragma unroll
  for (int j = 0; j < 16; j++) s[j] = P2_ADD3(s[j], sums[j & 3]);
}
// x^3: the square is left in (0, 2p) (no correction), which the second product tolerates
KB_HD Fp p2_sbox(Fp x) { return fp_raw(mont_mul_raw(mont_reduce_lazy((u64)x.v * x.v), x.v)); }

KB_HD void p2_internal_linear(Fp* s, const u32* diag, u32 big) {
  Fp sum = P2_ADD3(P2_ADD2(P2_ADD1(P2_ADD1(s[0], s[1]), P2_ADD1(s[2], s[3])), P2_ADD1(P2_ADD1(s[4], s[5]), P2_ADD1(s[6], s[7]))),
                   P2_ADD2(P2_ADD1(P2_ADD1(s[8], s[9]), P2_ADD1(s[10], s[11])), P2_ADD1(P2_ADD1(s[12], s[13]), P2_ADD1(s[14], s[15]))));
  // diag = [-2, 1, 2, 1/2, 3, 4, -1/2, -3, -4, 2^-8, 1/8, 2^-24, -2^-8, -1/8, -1/16, -2^-24]
  s[0] = sum - fp_double(s[0]);
  s[1] = sum + s[1];
  s[2] = sum + fp_double(s[2]);
  s[3] = sum + fp_halve(s[3]);
  s[4] = sum + fp_mul3(s[4]);
  s[5] = sum + fp_double(fp_double(s[5]));
  s[6] = sum - fp_halve(s[6]);
  s[7] = sum - fp_mul3(s[7]);
  s[8] = sum - fp_double(fp_double(s[8]));
#pragma unroll
  for (int i = 9; i < 16; i++) s[i] = sum + s[i] * fp_raw(diag[i]);
}

// generic permutation over a constants table reachable from the calling side
KB_HD void p2_permute_with(Fp* s, const P2Consts& C) {
  const u32 big = C.big;
  p2_external_linear(s, big);
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = p2_sbox(s[i] + fp_raw(C.ext[r][i]));
    p2_external_linear(s, big);
  }
#pragma unroll 1
  for (int r = 0; r < 13; r++) {
    s[0] = p2_sbox(s[0] + fp_raw(C.in[r]));
    p2_internal_linear(s, C.diag, big);
  }
#pragma unroll 1
  for (int r = 4; r < 8; r++) {
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = p2_sbox(s[i] + fp_raw(C.ext[r][i]));
    p2_external_linear(s, big);
  }
}

inline void p2_permute_host(Fp* s) { p2_permute_with(s, p2_host_consts()); }

}  // namespace zkb

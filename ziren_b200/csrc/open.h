// K4a/K4b: opening at zeta (barycentric evaluation over the low coset) and the reduced
// openings (DEEP quotients) that feed FRI.  Replaces the first half of [P3-upstream]
// TwoAdicFriPcs::open as called from crates/stark/src/prover.rs:546-556; the values are pinned
// by the verifier restatement crates/recursion/circuit/src/fri.rs:97-209.
#pragma once
#include "common.h"
#include "ntt.h"

namespace zkb {
// Barycentric weights for p(z) from evaluations on GENERATOR*H_n stored in bit-reversed order:
// out[c*n + pos], c<4 (EF component-major).
void bary_weights(const NttTables& tb, unsigned log_n, const Ef& z, u32* out, cudaStream_t s);
// ys[pt][col] = sum_pos w_pt[pos] * lde[col][pos], pos < n.  out: [npoints][W] EF (4 words each).
void eval_columns(const u32* lde, size_t H, size_t n, size_t W, const u32* w0, const u32* w1, int npoints, u32* out,
                  cudaStream_t s);
// out[c*H + pos] = 1 / (z - GENERATOR * w_H^bitrev(pos))
void inv_denominators(const NttTables& tb, unsigned log_h, const Ef& z, u32* out, cudaStream_t s);
// out[j] = alpha^j (4 words each), j < count
void ef_powers(const Ef& alpha, size_t count, u32* out, cudaStream_t s);
// ro[x] += sum_pt apow_off[pt] * (sum_j alpha^j (ys_pt[j] - p_j(x))) * invden_pt[x]
void reduce_matrix(const u32* lde, size_t H, size_t W, const u32* apow, const u32* ys, int npoints, const Ef& off0,
                   const Ef& off1, const u32* invden0, const u32* invden1, u32* ro, cudaStream_t s);
}  // namespace zkb

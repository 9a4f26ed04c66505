// ORACLE (test infrastructure, not product code).  Poseidon2-KoalaBear width 16, the
// padding-free sponge, the 2-to-1 truncated-permutation compression and the duplex challenger.
//
// Follows:
//   permutation ....... round schedule crates/primitives/src/lib.rs:1107-1123 (ROUNDS_F=8,
//                       ROUNDS_P=13; RC rows 0..3 / 4..16 (col 0) / 17..20); linear layers
//                       crates/core/machine/src/operations/poseidon2/air.rs:12-72 and
//                       crates/recursion/core/include/poseidon2.hpp:20-72; S-box x^3
//                       (crates/recursion/core/include/poseidon2_skinny.hpp:10-51).
//                       [P3-upstream Poseidon2::permute_mut applies the external linear layer
//                       once before the first round.]
//   sponge / compress . crates/recursion/circuit/src/hash.rs:40-49, :75-80
//   challenger ........ crates/recursion/circuit/src/challenger.rs:60-233
// Pinned by the reference's known-answer test examples/poseidon2/host/src/main.rs:33-37 and by
// the reference's own C++ headers compiled into oracle/_ref (see oracle/ref_build/).
#pragma once
#include "kb.h"
#include <cstring>

namespace zko {

static const u32 P2_RC[30][16] = {
#include "p2_rc.inc"
};

// diag of the internal matrix minus identity, canonical values
// (crates/core/machine/src/operations/poseidon2/air.rs:12-29).
static inline const F* p2_diag() {
  static F d[16];
  static bool init = false;
  if (!init) {
    const u32 raw[16] = {P - 2, 1, 2, (P + 1) >> 1, 3, 4, (P - 1) >> 1, P - 3, P - 4,
                         P - ((P - 1) >> 8), P - ((P - 1) >> 3), P - 127, (P - 1) >> 8,
                         (P - 1) >> 3, (P - 1) >> 4, 127};
    for (int i = 0; i < 16; i++) d[i] = F(raw[i]);
    init = true;
  }
  return d;
}

static inline void p2_m4(F* x) {
  F t01 = x[0] + x[1], t23 = x[2] + x[3];
  F t0123 = t01 + t23;
  F t01123 = t0123 + x[1], t01233 = t0123 + x[3];
  F x0d = x[0] + x[0], x2d = x[2] + x[2];
  x[3] = t01233 + x0d;
  x[1] = t01123 + x2d;
  x[0] = t01123 + t01;
  x[2] = t01233 + t23;
}
static inline void p2_external_linear(F* s) {
  for (int j = 0; j < 16; j += 4) p2_m4(s + j);
  F sums[4];
  for (int k = 0; k < 4; k++) sums[k] = s[k] + s[4 + k] + s[8 + k] + s[12 + k];
  for (int j = 0; j < 16; j++) s[j] += sums[j & 3];
}
static inline void p2_internal_linear(F* s) {
  const F* d = p2_diag();
  F sum = F::zero();
  for (int i = 0; i < 16; i++) sum += s[i];
  for (int i = 0; i < 16; i++) s[i] = s[i] * d[i] + sum;
}
static inline F p2_sbox(F x) { return x * x * x; }

static inline void poseidon2_permute(F* s) {
  p2_external_linear(s);
  for (int r = 0; r < 4; r++) {
    for (int i = 0; i < 16; i++) s[i] = p2_sbox(s[i] + F(P2_RC[r][i]));
    p2_external_linear(s);
  }
  for (int r = 0; r < 13; r++) {
    s[0] = p2_sbox(s[0] + F(P2_RC[4 + r][0]));
    p2_internal_linear(s);
  }
  for (int r = 0; r < 4; r++) {
    for (int i = 0; i < 16; i++) s[i] = p2_sbox(s[i] + F(P2_RC[17 + r][i]));
    p2_external_linear(s);
  }
}

typedef std::array<F, 8> Digest;

// PaddingFreeSponge<Perm,16,8,8>: overwrite-mode absorb, no padding.
static inline Digest sponge_hash(const F* in, size_t n) {
  F st[16];
  for (size_t off = 0; off < n; off += 8) {
    size_t len = n - off < 8 ? n - off : 8;
    for (size_t i = 0; i < len; i++) st[i] = in[off + i];
    poseidon2_permute(st);
  }
  Digest d;
  for (int i = 0; i < 8; i++) d[i] = st[i];
  return d;
}
// TruncatedPermutation<Perm,2,8,16>
static inline Digest compress2(const Digest& l, const Digest& r) {
  F st[16];
  for (int i = 0; i < 8; i++) { st[i] = l[i]; st[8 + i] = r[i]; }
  poseidon2_permute(st);
  Digest d;
  for (int i = 0; i < 8; i++) d[i] = st[i];
  return d;
}

// DuplexChallenger<F, Perm, 16, 8>
struct Challenger {
  F state[16];
  std::vector<F> in_buf, out_buf;

  void duplexing() {
    assert(in_buf.size() <= 8);
    for (size_t i = 0; i < in_buf.size(); i++) state[i] = in_buf[i];
    in_buf.clear();
    poseidon2_permute(state);
    out_buf.assign(state, state + 8);
  }
  void observe(F v) {
    out_buf.clear();
    in_buf.push_back(v);
    if (in_buf.size() == 8) duplexing();
  }
  void observe_slice(const F* v, size_t n) { for (size_t i = 0; i < n; i++) observe(v[i]); }
  void observe_digest(const Digest& d) { observe_slice(d.data(), 8); }
  void observe_ext(const E& e) { observe_slice(e.c, 4); }
  F sample() {
    if (!in_buf.empty() || out_buf.empty()) duplexing();
    F r = out_buf.back();
    out_buf.pop_back();
    return r;
  }
  E sample_ext() {
    E e;
    for (int i = 0; i < 4; i++) e.c[i] = sample();
    return e;
  }
  u32 sample_bits(unsigned bits) { return sample().v & ((1u << bits) - 1); }
  bool check_witness(unsigned bits, F w) { observe(w); return sample_bits(bits) == 0; }
  // [P3-upstream GrindingChallenger::grind] searches with rayon find_any (any valid witness);
  // the oracle takes the smallest valid witness so that proofs are deterministic.
  F grind(unsigned bits) {
    for (u32 w = 0; w < P; w++) {
      Challenger c = *this;
      if (c.check_witness(bits, F(w))) { bool ok = check_witness(bits, F(w)); assert(ok); (void)ok; return F(w); }
    }
    assert(false);
    return F(0);
  }
  // flat (de)serialisation: state[16], n_in, in[8], n_out, out[8]  (34 words)
  void to_words(u32* w) const {
    for (int i = 0; i < 16; i++) w[i] = state[i].v;
    w[16] = (u32)in_buf.size();
    for (int i = 0; i < 8; i++) w[17 + i] = i < (int)in_buf.size() ? in_buf[i].v : 0;
    w[25] = (u32)out_buf.size();
    for (int i = 0; i < 8; i++) w[26 + i] = i < (int)out_buf.size() ? out_buf[i].v : 0;
  }
  void from_words(const u32* w) {
    for (int i = 0; i < 16; i++) state[i] = F(w[i]);
    in_buf.clear(); out_buf.clear();
    for (u32 i = 0; i < w[16]; i++) in_buf.push_back(F(w[17 + i]));
    for (u32 i = 0; i < w[25]; i++) out_buf.push_back(F(w[26 + i]));
  }
};

}  // namespace zko

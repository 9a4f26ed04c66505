// ORACLE (test infrastructure, not product code).  CPU restatement of the KoalaBear field,
// its degree-4 binomial extension and the two-adic domain helpers used by the reference's
// STARK hot path.  Values are kept in CANONICAL form (plain residues 0..p-1), on purpose
// different from the Montgomery representation the CUDA path uses, so that representation
// bugs surface in parity tests.
//
// Follows:
//   field constants ........ crates/core/machine/include/kb31_t.hpp:27-34 (MOD = 0x7f000001)
//   EF4 = F[x]/(x^4 - 3) ... crates/stark/src/air/extension.rs:55-75 (W = 3)
//   two-adic generators .... [P3-upstream p3-koala-bear]; re-derived: g_24 = 3^127
//   domains / selectors .... crates/recursion/circuit/src/domain.rs:32-89
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>
#include <array>
#include <cassert>

namespace zko {

typedef uint32_t u32;
typedef uint64_t u64;

static const u32 P = 0x7f000001u;
static const u32 GENERATOR = 3;  // multiplicative generator, [P3-upstream]; 3^127 has order 2^24

struct F {
  u32 v;
  F() : v(0) {}
  explicit F(u32 x) : v(x) {}
  static F from_u64(u64 x) { return F((u32)(x % P)); }
  static F zero() { return F(0); }
  static F one() { return F(1); }
  bool operator==(const F& o) const { return v == o.v; }
  bool operator!=(const F& o) const { return v != o.v; }
  bool is_zero() const { return v == 0; }
};
static inline F operator+(F a, F b) { u32 s = a.v + b.v; return F(s >= P ? s - P : s); }
static inline F operator-(F a, F b) { return F(a.v >= b.v ? a.v - b.v : a.v + P - b.v); }
static inline F operator-(F a) { return F(a.v ? P - a.v : 0); }
static inline F operator*(F a, F b) { return F((u32)(((u64)a.v * b.v) % P)); }
static inline F& operator+=(F& a, F b) { a = a + b; return a; }
static inline F& operator-=(F& a, F b) { a = a - b; return a; }
static inline F& operator*=(F& a, F b) { a = a * b; return a; }

static inline F fpow(F b, u64 e) {
  F r = F::one();
  while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
  return r;
}
static inline F finv(F a) { assert(!a.is_zero()); return fpow(a, P - 2); }

// Element of order 2^bits.  Table entry 24 is 3^127; lower entries by repeated squaring.
static inline F two_adic_generator(unsigned bits) {
  assert(bits <= 24);
  F g = fpow(F(GENERATOR), 127);
  for (unsigned i = bits; i < 24; i++) g *= g;
  return g;
}

// ---- EF4 = F[x] / (x^4 - 3) ------------------------------------------------------------
struct E {
  F c[4];
  E() {}
  explicit E(F a) { c[0] = a; }
  E(F a, F b, F cc, F d) { c[0] = a; c[1] = b; c[2] = cc; c[3] = d; }
  static E zero() { return E(); }
  static E one() { return E(F::one()); }
  bool operator==(const E& o) const { return c[0]==o.c[0] && c[1]==o.c[1] && c[2]==o.c[2] && c[3]==o.c[3]; }
  bool operator!=(const E& o) const { return !(*this == o); }
  bool is_zero() const { return c[0].is_zero() && c[1].is_zero() && c[2].is_zero() && c[3].is_zero(); }
};
static inline E operator+(const E& a, const E& b) { return E(a.c[0]+b.c[0], a.c[1]+b.c[1], a.c[2]+b.c[2], a.c[3]+b.c[3]); }
static inline E operator-(const E& a, const E& b) { return E(a.c[0]-b.c[0], a.c[1]-b.c[1], a.c[2]-b.c[2], a.c[3]-b.c[3]); }
static inline E operator-(const E& a) { return E(-a.c[0], -a.c[1], -a.c[2], -a.c[3]); }
static inline E operator*(const E& a, F b) { return E(a.c[0]*b, a.c[1]*b, a.c[2]*b, a.c[3]*b); }
static inline E operator+(const E& a, F b) { E r = a; r.c[0] += b; return r; }
static inline E operator-(const E& a, F b) { E r = a; r.c[0] -= b; return r; }
static inline E operator*(const E& a, const E& b) {
  // schoolbook with x^4 = 3
  const F W(3);
  E r;
  r.c[0] = a.c[0]*b.c[0] + W*(a.c[1]*b.c[3] + a.c[2]*b.c[2] + a.c[3]*b.c[1]);
  r.c[1] = a.c[0]*b.c[1] + a.c[1]*b.c[0] + W*(a.c[2]*b.c[3] + a.c[3]*b.c[2]);
  r.c[2] = a.c[0]*b.c[2] + a.c[1]*b.c[1] + a.c[2]*b.c[0] + W*(a.c[3]*b.c[3]);
  r.c[3] = a.c[0]*b.c[3] + a.c[1]*b.c[2] + a.c[2]*b.c[1] + a.c[3]*b.c[0];
  return r;
}
static inline E& operator+=(E& a, const E& b) { a = a + b; return a; }
static inline E& operator-=(E& a, const E& b) { a = a - b; return a; }
static inline E& operator*=(E& a, const E& b) { a = a * b; return a; }

static inline E epow(E b, u64 e) {
  E r = E::one();
  while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
  return r;
}
// Inverse through the norm to the quadratic subfield F[y]/(y^2-3), y = x^2:
// a = A + x B with A = a0 + a2 y, B = a1 + a3 y;  a^-1 = (A - x B) / (A^2 - y B^2).
static inline E einv(const E& a) {
  assert(!a.is_zero());
  const F W(3);
  // A^2 = (a0^2 + 3 a2^2) + (2 a0 a2) y ;  B^2 = (a1^2 + 3 a3^2) + (2 a1 a3) y
  F A2_0 = a.c[0]*a.c[0] + W*a.c[2]*a.c[2], A2_1 = (a.c[0]*a.c[2]) + (a.c[0]*a.c[2]);
  F B2_0 = a.c[1]*a.c[1] + W*a.c[3]*a.c[3], B2_1 = (a.c[1]*a.c[3]) + (a.c[1]*a.c[3]);
  // y * B^2 = 3 B2_1 + B2_0 y
  F N0 = A2_0 - W*B2_1, N1 = A2_1 - B2_0;        // N = N0 + N1 y in the quadratic subfield
  // N^-1 = (N0 - N1 y) / (N0^2 - 3 N1^2)
  F d = finv(N0*N0 - W*N1*N1);
  F I0 = N0*d, I1 = -(N1*d);
  // (A - xB) * (I0 + I1 y):  A*I = (a0 I0 + 3 a2 I1) + (a0 I1 + a2 I0) y ; same for B
  F AI0 = a.c[0]*I0 + W*a.c[2]*I1, AI1 = a.c[0]*I1 + a.c[2]*I0;
  F BI0 = a.c[1]*I0 + W*a.c[3]*I1, BI1 = a.c[1]*I1 + a.c[3]*I0;
  return E(AI0, -BI0, AI1, -BI1);
}

static inline unsigned log2_strict(size_t n) {
  unsigned l = 0;
  while (((size_t)1 << l) < n) l++;
  assert(((size_t)1 << l) == n);
  return l;
}
static inline size_t bitrev(size_t x, unsigned bits) {
  size_t r = 0;
  for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}

// Lagrange selectors of the trace domain H_n (shift 1) at an arbitrary EF point.
// crates/recursion/circuit/src/domain.rs:46-64.
struct SelectorsE { E is_first_row, is_last_row, is_transition, inv_zeroifier; };
static inline SelectorsE selectors_at_point(unsigned log_n, const E& z) {
  E zh = z;
  for (unsigned i = 0; i < log_n; i++) zh *= zh;
  zh = zh - F::one();
  F ginv = finv(two_adic_generator(log_n));
  SelectorsE s;
  s.is_first_row = zh * einv(z - F::one());
  s.is_last_row = zh * einv(z - ginv);
  s.is_transition = z - ginv;
  s.inv_zeroifier = einv(zh);
  return s;
}

}  // namespace zko

// CPU restatement of the Global chip's trace generation - TEST INFRASTRUCTURE, the checker for
// ziren_b200/csrc/tracegen_global.cuh, never the thing shipped or measured.  Canonical residues, the reference's own
// formulation step by step:
//   SepticExtension  crates/stark/src/septic_extension.rs: Mul :306-323, frobenius / double_frobenius :581-605, pow_r_1 :607-613,
//                    inv :615-619, is_square :621-628, sqrt with Cipolla's algorithm :632-680 and :708-740
//   SepticCurve      crates/stark/src/septic_curve.rs: add_incomplete :48-53, double :62-78, curve_formula :97-121, lift_x
//                    :130-154, sum_checker_x :159-166, SepticCurveComplete::add :198-217
//   GlobalChip       crates/core/machine/src/global/mod.rs:115-194; GlobalLookupOperation::populate
//                    crates/core/machine/src/operations/global_lookup.rs:31-90; GlobalAccumulationOperation::populate_real /
//                    populate_dummy crates/core/machine/src/operations/global_accumulation.rs:84-127
// The cumulative sum is the reference's left-to-right scan.  The Frobenius constants z^(i p), z^(i p^2) are computed by
// exponentiation when first used (the reference writes them out, septic_extension.rs:429-578; tests compare them with the
// reference's C++ twin, crates/core/machine/include/kb31_septic_extension_t.hpp, when oracle/_ref is built).
#pragma once
#include <stdexcept>
#include <vector>
#include "kb.h"

namespace zko {

struct Septic {
  F c[7];
  bool operator==(const Septic& o) const { for (int i = 0; i < 7; i++) if (c[i] != o.c[i]) return false; return true; }
  bool is_zero() const { for (int i = 0; i < 7; i++) if (!c[i].is_zero()) return false; return true; }
};
static inline Septic septic_from_base(F a) { Septic r; r.c[0] = a; return r; }
static inline Septic operator+(const Septic& a, const Septic& b) { Septic r; for (int i = 0; i < 7; i++) r.c[i] = a.c[i] + b.c[i]; return r; }
static inline Septic operator-(const Septic& a, const Septic& b) { Septic r; for (int i = 0; i < 7; i++) r.c[i] = a.c[i] - b.c[i]; return r; }
static inline Septic operator-(const Septic& a) { Septic r; for (int i = 0; i < 7; i++) r.c[i] = -a.c[i]; return r; }
static inline Septic operator*(const Septic& a, F b) { Septic r; for (int i = 0; i < 7; i++) r.c[i] = a.c[i] * b; return r; }
static inline Septic operator*(const Septic& a, const Septic& b) {
  F res[13];
  for (int i = 0; i < 7; i++)
    for (int j = 0; j < 7; j++) res[i + j] += a.c[i] * b.c[j];
  Septic r;
  for (int i = 0; i < 7; i++) r.c[i] = res[i];
  for (int i = 7; i < 13; i++) {
    r.c[i - 7] += res[i] * F(8);
    r.c[i - 6] -= res[i] * F(2);
  }
  return r;
}
static inline Septic septic_pow(Septic b, u64 e) {
  Septic r = septic_from_base(F::one());
  while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; }
  return r;
}
struct FrobeniusTables { Septic zp[7], zp2[7]; };
static inline const FrobeniusTables& frobenius_tables() {
  static const FrobeniusTables t = [] {
    FrobeniusTables k;
    Septic z;
    z.c[1] = F::one();
    const Septic zp = septic_pow(z, P), zp2 = septic_pow(zp, P);
    k.zp[0] = k.zp2[0] = septic_from_base(F::one());
    for (int i = 1; i < 7; i++) { k.zp[i] = k.zp[i - 1] * zp; k.zp2[i] = k.zp2[i - 1] * zp2; }
    return k;
  }();
  return t;
}
static inline Septic septic_frobenius(const Septic& a) {
  Septic r;
  for (int i = 0; i < 7; i++) r = r + frobenius_tables().zp[i] * a.c[i];
  return r;
}
static inline Septic septic_double_frobenius(const Septic& a) {
  Septic r;
  for (int i = 0; i < 7; i++) r = r + frobenius_tables().zp2[i] * a.c[i];
  return r;
}
static inline Septic septic_pow_r_1(const Septic& a) {
  const Septic base = septic_frobenius(a) * septic_double_frobenius(a);
  const Septic base_p2 = septic_double_frobenius(base);
  const Septic base_p4 = septic_double_frobenius(base_p2);
  return base * base_p2 * base_p4;
}
static inline Septic septic_inv(const Septic& a) {
  const Septic pow_r_1 = septic_pow_r_1(a);
  const Septic pow_r = pow_r_1 * a;
  for (int i = 1; i < 7; i++) if (!pow_r.c[i].is_zero()) throw std::runtime_error("oracle: the norm is not in the base field");
  return pow_r_1 * finv(pow_r.c[0]);
}
// SepticExtension::sqrt: false when n is not a square
static inline bool septic_sqrt(const Septic& n, Septic& out) {
  if (n.is_zero() || n == septic_from_base(F::one())) { out = n; return true; }
  const Septic pow_r = septic_pow_r_1(n) * n;
  const F numerator = pow_r.c[0];
  if (fpow(numerator, (P - 1) / 2) != F::one()) return false;
  Septic n_iter = n, n_power = n;
  for (int i = 1; i < 30; i++) {
    n_iter = n_iter * n_iter;
    if (i >= 23) n_power = n_power * n_iter;
  }
  Septic n_frobenius = septic_frobenius(n_power);
  Septic denominator = n_frobenius;
  n_frobenius = septic_double_frobenius(n_frobenius);
  denominator = denominator * n_frobenius;
  n_frobenius = septic_double_frobenius(n_frobenius);
  denominator = denominator * n_frobenius;
  denominator = denominator * n;
  // Cipolla: a with a^2 - base a non-residue, then (a + sqrt(a^2 - base))^((p + 1) / 2)
  const F base = finv(numerator);
  F a = F::one(), nonresidue = F::one() - base;
  while (fpow(nonresidue, (P - 1) / 2) == F::one()) {
    a *= F(GENERATOR);
    nonresidue = a * a - base;
  }
  F re = F::one(), im = F::zero(), bre = a, bim = F::one();
  for (u64 e = ((u64)P + 1) / 2; e; e >>= 1) {
    if (e & 1) { const F t = re * bre + nonresidue * im * bim; im = re * bim + im * bre; re = t; }
    const F t = bre * bre + nonresidue * bim * bim; bim = bre * bim + bim * bre; bre = t;
  }
  out = denominator * re;
  return true;
}
static inline Septic curve_formula(const Septic& x) {
  Septic three_z, three;
  three_z.c[1] = F(3);
  three.c[0] = F(3);
  return x * x * x + x * three_z - three;
}
struct CurvePoint { Septic x, y; bool infinity = false; };
// SepticCurve::lift_x: the point with 1 <= y[6] <= (p - 1) / 2 and the offset used
static inline CurvePoint curve_lift_x(const Septic& m, u32& offset) {
  for (u32 o = 0; o < 256; o++) {
    Septic x = m;
    x.c[6] = m.c[6] * F(256) + F(o);
    Septic y;
    if (!septic_sqrt(curve_formula(x), y)) continue;
    if (y.c[6].is_zero()) continue;                        // is_exception
    if (y.c[6].v >= (P + 1) / 2) y = -y;                   // is_send
    offset = o;
    CurvePoint p;
    p.x = x; p.y = y;
    return p;
  }
  throw std::runtime_error("oracle: curve point couldn't be found after 256 attempts");
}
static inline CurvePoint curve_add_incomplete(const CurvePoint& a, const CurvePoint& b) {
  const Septic slope = (b.y - a.y) * septic_inv(b.x - a.x);
  CurvePoint r;
  r.x = slope * slope - a.x - b.x;
  r.y = slope * (a.x - r.x) - a.y;
  return r;
}
static inline CurvePoint curve_double(const CurvePoint& a) {
  Septic three_z;
  three_z.c[1] = F(3);
  const Septic slope = (a.x * a.x * F(3) + three_z) * septic_inv(a.y * F(2));
  CurvePoint r;
  r.x = slope * slope - a.x * F(2);
  r.y = slope * (a.x - r.x) - a.y;
  return r;
}
// SepticCurveComplete::add
static inline CurvePoint curve_add_complete(const CurvePoint& a, const CurvePoint& b) {
  if (a.infinity) return b;
  if (b.infinity) return a;
  if (!(a.x == b.x)) return curve_add_incomplete(a, b);
  if (a.y == b.y) return curve_double(a);
  CurvePoint inf;
  inf.infinity = true;
  return inf;
}
static inline Septic curve_sum_checker_x(const CurvePoint& p1, const CurvePoint& p2, const CurvePoint& p3) {
  const Septic dx = p2.x - p1.x, dy = p2.y - p1.y;
  return (p1.x + p2.x + p3.x) * (dx * dx) - dy * dy;
}
static const u32 CURVE_CUMULATIVE_SUM_START[14] = {637514027, 1595065213, 1998064738, 72333738, 1211544370, 822986770, 1518535784,
                                                   1604177449, 90440090, 259343427, 140470264, 1162099742, 941559812, 1064053343};
static const u32 CURVE_WITNESS_DUMMY_POINT[14] = {1706420302, 1319108093, 148224806, 26874985, 1766171812, 1645633948, 2028659224,
                                                  942390502, 1239997438, 458866455, 1843332012, 1309764648, 572807436, 74267719};
static inline CurvePoint curve_point_from_words(const u32* w) {
  CurvePoint p;
  for (int i = 0; i < 7; i++) { p.x.c[i] = F(w[i]); p.y.c[i] = F(w[7 + i]); }
  return p;
}

// events: n GlobalLookupEvent records of 8 words {message[7], is_receive | kind << 8}
// (crates/core/executor/src/events/global.rs:6-15); out: height x 99 canonical words, row-major
enum { GLOBAL_WIDTH = 99, GLOBAL_EVENT_WORDS = 8 };
static inline void global_trace(const u32* ev, size_t n, size_t height, u32* out) {
  if (n > height) throw std::runtime_error("oracle: more events than rows");
  std::vector<CurvePoint> sums;
  sums.push_back(curve_point_from_words(CURVE_CUMULATIVE_SUM_START));
  for (size_t i = 0; i < n; i++) {
    const u32* e = ev + GLOBAL_EVENT_WORDS * i;
    u32* row = out + i * GLOBAL_WIDTH;
    const bool is_receive = (e[7] & 0xff) != 0;
    const u32 kind = (e[7] >> 8) & 0xff;
    // GlobalLookupOperation::get_digest
    Septic x_start;
    for (int k = 0; k < 7; k++) x_start.c[k] = F(e[k] % P);
    x_start.c[0] += F(kind << 16);
    u32 offset = 0;
    CurvePoint point = curve_lift_x(x_start, offset);
    if (!is_receive) point.y = -point.y;
    int at = 0;
    for (int k = 0; k < 7; k++) row[at++] = e[k] % P;
    row[at++] = kind;
    for (int k = 0; k < 8; k++) row[at++] = (offset >> k) & 1;
    for (int k = 0; k < 7; k++) row[at++] = point.x.c[k].v;
    for (int k = 0; k < 7; k++) row[at++] = point.y.c[k].v;
    const u32 range_check_value = is_receive ? point.y.c[6].v - 1 : point.y.c[6].v - (P + 1) / 2;
    F top_7_bits = F::zero();
    for (int k = 0; k < 30; k++) {
      row[at++] = (range_check_value >> k) & 1;
      if (k >= 23) top_7_bits += F((range_check_value >> k) & 1);
    }
    top_7_bits -= F(7);
    row[at++] = finv(top_7_bits).v;
    row[at++] = is_receive ? 1 : 0; row[at++] = is_receive ? 0 : 1; row[at++] = 1;
    if (at != 64) throw std::runtime_error("oracle: Global lookup columns mismatch");
    sums.push_back(curve_add_complete(sums.back(), point));
    if (sums.back().infinity) throw std::runtime_error("oracle: point() called for point at infinity");
  }
  // no event at all: the reference's scan is empty and the final digest is the dummy point (global/mod.rs:162-165)
  const CurvePoint dummy = curve_point_from_words(CURVE_WITNESS_DUMMY_POINT), final_digest = n ? sums.back() : dummy;
  const Septic final_sum_checker = curve_sum_checker_x(final_digest, dummy, final_digest);
  for (size_t i = 0; i < height; i++) {
    u32* row = out + i * GLOBAL_WIDTH;
    const CurvePoint& first = i < n ? sums[i] : final_digest;
    const CurvePoint& second = i < n ? sums[i + 1] : final_digest;
    if (i >= n) {                                          // populate_dummy of the lookup columns; the rest of the row is zero
      for (int k = 0; k < 64; k++) row[k] = 0;
      for (int k = 0; k < 14; k++) row[16 + k] = CURVE_WITNESS_DUMMY_POINT[k];
    }
    for (int k = 0; k < 7; k++) {
      row[64 + k] = first.x.c[k].v; row[71 + k] = first.y.c[k].v;
      row[78 + k] = i < n ? 0 : final_sum_checker.c[k].v;
      row[85 + k] = second.x.c[k].v; row[92 + k] = second.y.c[k].v;
    }
  }
}

}  // namespace zko

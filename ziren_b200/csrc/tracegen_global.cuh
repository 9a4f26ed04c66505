// K6c: trace generation of the Global chip, the one core table whose rows are compute-heavy: every
// global lookup is lifted to a point of the septic curve y^2 = x^3 + 3z x - 3 over F_p^7 = F_p[z] / (z^7 + 2z - 8)
// (crates/stark/src/septic_curve.rs, septic_extension.rs) and the points are summed along the table.
//   GlobalChip::generate_trace            crates/core/machine/src/global/mod.rs:115-194 (GlobalCols :53-63, 99 columns)
//   GlobalLookupOperation::populate       crates/core/machine/src/operations/global_lookup.rs:31-90
//   GlobalAccumulationOperation           crates/core/machine/src/operations/global_accumulation.rs:84-127
//   SepticCurve::lift_x                   crates/stark/src/septic_curve.rs:130-154
// Three steps, each one thread per unit, written host/device so that the CPU suite walks the same code (tests/hostcheck):
//   lift     event -> curve point (a few thousand field products: square test through the norm, square root) and the row's
//            message / lookup columns
//   scan     inclusive scan of the points under the complete curve addition, in chunks: chunk totals, the same scan over the
//            totals, then each chunk again from its prefix.  Curve addition is associative, so the sums are those of the
//            reference's left-to-right scan.
//   finish   the accumulation columns of every row (two neighbouring sums), the dummy rows past the last event
// Differences from the reference's formulation that cannot be seen in the rows: the base-field square root is Tonelli-Shanks
// instead of Cipolla (lift_x fixes the sign of y afterwards), the Frobenius constants z^(ip), z^(ip^2) are computed at start-up
// by exponentiation instead of being written out.
#pragma once
#include "kb31.cuh"

namespace zkb {

constexpr int GLOBAL_WIDTH = 99, GLOBAL_EVENT_WORDS = 8, GLOBAL_POINT_WORDS = 14, GLOBAL_SCAN_CHUNK = 32;
// canonical: CURVE_CUMULATIVE_SUM_START_{X,Y} (crates/stark/src/septic_digest.rs:9-14), CURVE_WITNESS_DUMMY_POINT_{X,Y}
// (crates/stark/src/septic_curve.rs:14-19)
constexpr u32 GLOBAL_START_POINT[14] = {637514027, 1595065213, 1998064738, 72333738, 1211544370, 822986770, 1518535784,
                                        1604177449, 90440090, 259343427, 140470264, 1162099742, 941559812, 1064053343};
constexpr u32 GLOBAL_DUMMY_POINT[14] = {1706420302, 1319108093, 148224806, 26874985, 1766171812, 1645633948, 2028659224,
                                        942390502, 1239997438, 458866455, 1843332012, 1309764648, 572807436, 74267719};

struct Sep { Fp c[7]; };
KB_HD Sep sep_zero() { Sep r; for (int i = 0; i < 7; i++) r.c[i] = fp_zero(); return r; }
KB_HD Sep sep_from_base(Fp a) { Sep r = sep_zero(); r.c[0] = a; return r; }
KB_HD Sep operator+(const Sep& a, const Sep& b) { Sep r; for (int i = 0; i < 7; i++) r.c[i] = a.c[i] + b.c[i]; return r; }
KB_HD Sep operator-(const Sep& a, const Sep& b) { Sep r; for (int i = 0; i < 7; i++) r.c[i] = a.c[i] - b.c[i]; return r; }
KB_HD Sep operator-(const Sep& a) { Sep r; for (int i = 0; i < 7; i++) r.c[i] = -a.c[i]; return r; }
KB_HD Sep operator*(const Sep& a, Fp b) { Sep r; for (int i = 0; i < 7; i++) r.c[i] = a.c[i] * b; return r; }
KB_HD bool sep_is_zero(const Sep& a) { u32 o = 0; for (int i = 0; i < 7; i++) o |= a.c[i].v; return o == 0; }
KB_HD bool operator==(const Sep& a, const Sep& b) { u32 o = 0; for (int i = 0; i < 7; i++) o |= a.c[i].v ^ b.c[i].v; return o == 0; }
// schoolbook product, then z^(7+k) = 8 z^k - 2 z^(k+1) (septic_extension.rs:306-323)
KB_HD Sep operator*(const Sep& a, const Sep& b) {
  Fp t[13];
  for (int k = 0; k < 13; k++) {
    // up to seven products of residues below p: accumulate in 64 bits, reduce once (7 p^2 < 2^65 does not fit: split 4 + 3)
    u64 lo = 0, hi = 0;
    for (int i = 0; i < 7; i++) {
      const int j = k - i;
      if (j < 0 || j > 6) continue;
      if (i < 4) lo += mul_wide(a.c[i].v, b.c[j].v); else hi += mul_wide(a.c[i].v, b.c[j].v);
    }
    t[k] = fp_raw(mont_reduce_wide(lo)) + fp_raw(mont_reduce_wide(hi));
  }
  Sep r;
  for (int i = 0; i < 7; i++) r.c[i] = t[i];
  const Fp eight = fp_from_canonical(8), two = fp_from_canonical(2);
  for (int k = 7; k < 13; k++) {
    r.c[k - 7] += t[k] * eight;
    r.c[k - 6] -= t[k] * two;
  }
  return r;
}
KB_HD Sep sep_sqr(const Sep& a) { return a * a; }

// SepticCurveComplete (septic_curve.rs:190-235): the point at infinity is stored as x = y = 0, which is not on the curve
struct CurvePt { Sep x, y; };

// z^(i p) and z^(i p^2) for i = 1..6 (septic_extension.rs z_pow_p, z_pow_p2), computed by global_build_consts
struct GlobalConsts {
  Sep zp[6], zp2[6];
  CurvePt start, dummy;  // SepticDigest::zero().0 and SepticCurve::dummy(), Montgomery
  Fp ts_root;          // 3^127: a generator of the 2^24-th roots of unity, for Tonelli-Shanks
  Fp neg_inv[8];       // -1/k for k = 1..7 (range_check_witness: the inverse of (number of set top bits) - 7)
};
KB_HD Sep sep_pow(Sep b, u64 e) {
  Sep r = sep_from_base(fp_one());
  while (e) { if (e & 1) r = r * b; b = b * b; e >>= 1; }
  return r;
}
inline void global_build_consts(GlobalConsts& k) {
  Sep z = sep_zero();
  z.c[1] = fp_one();
  const Sep zp = sep_pow(z, KB_P), zp2 = sep_pow(zp, KB_P);
  k.zp[0] = zp; k.zp2[0] = zp2;
  for (int i = 1; i < 6; i++) { k.zp[i] = k.zp[i - 1] * zp; k.zp2[i] = k.zp2[i - 1] * zp2; }
  for (int i = 0; i < 7; i++) {
    k.start.x.c[i] = fp_from_canonical(GLOBAL_START_POINT[i]); k.start.y.c[i] = fp_from_canonical(GLOBAL_START_POINT[7 + i]);
    k.dummy.x.c[i] = fp_from_canonical(GLOBAL_DUMMY_POINT[i]); k.dummy.y.c[i] = fp_from_canonical(GLOBAL_DUMMY_POINT[7 + i]);
  }
  k.ts_root = fp_pow(fp_from_canonical(KB_GEN), 127);
  k.neg_inv[0] = fp_zero();
  for (u32 i = 1; i < 8; i++) k.neg_inv[i] = -fp_inv(fp_from_canonical(i));
}
// a^p and a^(p^2): coefficient-wise against the constants (septic_extension.rs:581-605)
KB_HD Sep sep_frobenius(const Sep& a, const GlobalConsts& k) {
  Sep r = sep_from_base(a.c[0]);
  for (int i = 0; i < 6; i++) r = r + k.zp[i] * a.c[i + 1];
  return r;
}
KB_HD Sep sep_double_frobenius(const Sep& a, const GlobalConsts& k) {
  Sep r = sep_from_base(a.c[0]);
  for (int i = 0; i < 6; i++) r = r + k.zp2[i] * a.c[i + 1];
  return r;
}
// a^(p + p^2 + ... + p^6): a times this is the norm, an element of F_p (septic_extension.rs:607-613)
KB_HD Sep sep_pow_r_1(const Sep& a, const GlobalConsts& k) {
  const Sep base = sep_frobenius(a, k) * sep_double_frobenius(a, k);
  const Sep base_p2 = sep_double_frobenius(base, k);
  const Sep base_p4 = sep_double_frobenius(base_p2, k);
  return base * base_p2 * base_p4;
}
// coefficient 0 of pow_r_1 * a, the norm (the other coefficients vanish): the constant term of the schoolbook product plus
// eight times its z^7 term (z^(7+k) = 8 z^k - 2 z^(k+1) has a constant term for k = 0 only)
KB_HD Fp sep_norm_with(const Sep& a, const Sep& pow_r_1) {
  Fp t7 = fp_zero();
  for (int i = 1; i < 7; i++) t7 += pow_r_1.c[i] * a.c[7 - i];
  return pow_r_1.c[0] * a.c[0] + t7 * fp_from_canonical(8);
}
KB_HD Sep sep_inv(const Sep& a, const GlobalConsts& k) {
  const Sep q = sep_pow_r_1(a, k);
  return q * fp_inv(sep_norm_with(a, q));
}
// square root in F_p (p - 1 = 2^24 * 127), Tonelli-Shanks; a must be a nonzero square
KB_HD Fp fp_sqrt_ts(Fp a, Fp root) {
  Fp x = fp_pow(a, 64);            // a^((127 + 1) / 2)
  Fp b = fp_pow(a, 127);           // x^2 = a b, b of order dividing 2^23
  Fp c = root;
  int m = 24;
  for (int round = 0; round < 24 && b != fp_one(); round++) {      // at most 23 rounds for a square; bounded for anything else
    int i = 0;
    Fp t = b;
    while (t != fp_one() && i < m) { t *= t; i++; }
    Fp g = c;
    for (int j = 0; j < m - i - 1; j++) g *= g;
    x *= g; c = g * g; b *= c; m = i;
  }
  return x;
}
// SepticExtension::is_square (septic_extension.rs:621-628): n is a square iff its norm n^(1 + p + ... + p^6) is one in F_p.
// Zero counts as a non-square here: lift_x skips y = 0 anyway.
KB_HD bool sep_is_square(const Sep& n, const GlobalConsts& k, Fp& norm) {
  norm = sep_norm_with(n, sep_pow_r_1(n, k));
  return fp_pow(norm, (KB_P - 1) / 2) == fp_one();
}
// SepticExtension::sqrt (septic_extension.rs:632-680) of a square with the given norm: n^((r + 1) / 2), r = 1 + p + ... + p^6,
// squares to norm(n) * n, so it is divided by a root of the norm.
KB_HD Sep sep_sqrt_of_square(const Sep& n, Fp norm, const GlobalConsts& k) {
  Sep it = n, pw = n;                              // n^((p + 1) / 2) = n^(1 + 2^23 + ... + 2^29)
  for (int i = 1; i < 30; i++) {
    it = sep_sqr(it);
    if (i >= 23) pw = pw * it;
  }
  Sep f = sep_frobenius(pw, k);
  Sep den = f;
  f = sep_double_frobenius(f, k); den = den * f;
  f = sep_double_frobenius(f, k); den = den * f;
  den = den * n;
  return den * fp_sqrt_ts(fp_inv(norm), k.ts_root);
}
KB_HD bool sep_sqrt(const Sep& n, const GlobalConsts& k, Sep& out) {
  Fp norm;
  if (!sep_is_square(n, k, norm)) return false;
  out = sep_sqrt_of_square(n, norm, k);
  return true;
}
// x^3 + 3z x - 3 (septic_curve.rs:97-121)
KB_HD Sep curve_formula(const Sep& x) {
  Sep three_z = sep_zero();
  three_z.c[1] = fp_from_canonical(3);
  Sep r = x * x * x + x * three_z;
  r.c[0] -= fp_from_canonical(3);
  return r;
}

KB_HD bool curve_is_infinity(const CurvePt& p) { return sep_is_zero(p.x) && sep_is_zero(p.y); }
KB_HD CurvePt curve_infinity() { CurvePt p; p.x = sep_zero(); p.y = sep_zero(); return p; }
KB_HD CurvePt curve_add(const CurvePt& a, const CurvePt& b, const GlobalConsts& k) {
  if (curve_is_infinity(a)) return b;
  if (curve_is_infinity(b)) return a;
  Sep slope;
  if (!(a.x == b.x)) slope = (b.y - a.y) * sep_inv(b.x - a.x, k);            // add_incomplete, septic_curve.rs:48-53
  else if (a.y == b.y) {                                                      // double, septic_curve.rs:62-78
    Sep three_z = sep_zero();
    three_z.c[1] = fp_from_canonical(3);
    slope = (a.x * a.x * fp_from_canonical(3) + three_z) * sep_inv(a.y + a.y, k);
  } else return curve_infinity();
  CurvePt r;
  r.x = sep_sqr(slope) - a.x - b.x;
  r.y = slope * (a.x - r.x) - a.y;
  return r;
}
KB_HD CurvePt curve_load(const u32* w) {
  CurvePt p;
  for (int i = 0; i < 7; i++) { p.x.c[i] = fp_raw(w[i]); p.y.c[i] = fp_raw(w[7 + i]); }
  return p;
}
KB_HD void curve_store(u32* w, const CurvePt& p) {
  for (int i = 0; i < 7; i++) { w[i] = p.x.c[i].v; w[7 + i] = p.y.c[i].v; }
}
// SepticCurve::sum_checker_x (septic_curve.rs:159-166)
KB_HD Sep curve_sum_checker_x(const CurvePt& p1, const CurvePt& p2, const CurvePt& p3) {
  return (p1.x + p2.x + p3.x) * sep_sqr(p2.x - p1.x) - sep_sqr(p2.y - p1.y);
}

// where a row's column goes: row-major (rs = width, cs = 1) or column-major (rs = 1, cs = height)
struct GlobalOut {
  u32* p; size_t rs, cs;
  KB_HD void operator()(size_t row, int col, u32 v) const { p[row * rs + (size_t)col * cs] = v; }
};

// lift: columns 0..63 of row `row` from its event (GlobalLookupEvent, crates/core/executor/src/events/global.rs:6-15, as its
// eight #[repr(C)] words: message[7], then is_receive in byte 0 and kind in byte 1), and the row's point into points[1 + row]
// (points[0] is the start of the cumulative sum, SepticDigest::zero()).  A message that lifts nowhere in 256 offsets (the
// reference panics; probability 2^-256) leaves the point at infinity.
KB_HD void global_lift_row(const u32* e, size_t row, const GlobalConsts& k, const GlobalOut& out, u32* points) {
  const bool is_receive = (e[7] & 0xffu) != 0;
  const u32 kind = (e[7] >> 8) & 0xffu;
  Sep x;
  for (int i = 0; i < 7; i++) x.c[i] = fp_from_canonical(e[i]);
  x.c[0] += fp_from_canonical(kind << 16);
  const Fp m6 = x.c[6] * fp_from_canonical(256);
  CurvePt pt = curve_infinity();
  u32 offset = 0;
  // The search for the first offset whose curve value is a square (a norm and a Legendre symbol per trial, two trials on
  // average, a different number in every lane of a warp) is kept apart from the root (six times the work of a trial), so
  // that the lanes of a warp take their roots together instead of one trial position after the other.
  for (u32 o = 0; o < 256; o++) {
    Sep n;
    Fp norm;
    bool found = false;
    for (; o < 256; o++) {
      x.c[6] = m6 + fp_from_canonical(o);
      n = curve_formula(x);
      if (sep_is_square(n, k, norm)) { found = true; break; }
    }
    if (!found) break;
    const Sep y = sep_sqrt_of_square(n, norm, k);
    const u32 y6 = fp_to_canonical(y.c[6]);
    if (y6 == 0) continue;                                        // is_exception: on to the next offset
    // lift_x returns the root with 1 <= y6 <= (p - 1) / 2; a send takes the negated point (global_lookup.rs:31-43)
    const bool low = y6 <= (KB_P - 1) / 2;
    pt.x = x; pt.y = low == is_receive ? y : -y;
    offset = o;
    break;
  }
  for (int i = 0; i < 7; i++) out(row, i, fp_from_canonical(e[i]).v);
  out(row, 7, fp_from_canonical(kind).v);
  for (int i = 0; i < 8; i++) out(row, 8 + i, (offset >> i) & 1u ? KB_ONE : 0u);
  for (int i = 0; i < 7; i++) { out(row, 16 + i, pt.x.c[i].v); out(row, 23 + i, pt.y.c[i].v); }
  const u32 y6 = fp_to_canonical(pt.y.c[6]);
  const u32 range = is_receive ? y6 - 1u : y6 - (KB_P + 1) / 2;
  u32 top = 0;
  for (int i = 0; i < 30; i++) {
    out(row, 30 + i, (range >> i) & 1u ? KB_ONE : 0u);
    if (i >= 23) top += (range >> i) & 1u;
  }
  out(row, 60, k.neg_inv[(7u - top) & 7u].v);                       // 1 / (top - 7)
  out(row, 61, is_receive ? KB_ONE : 0u); out(row, 62, is_receive ? 0u : KB_ONE); out(row, 63, KB_ONE);
  curve_store(points + GLOBAL_POINT_WORDS * (row + 1), pt);
}

// scan, pass 1: the sum of chunk t of `pts` (n points of 14 words) into totals[t]
KB_HD void global_chunk_total(const u32* pts, size_t n, size_t t, const GlobalConsts& k, u32* totals) {
  const size_t lo = t * GLOBAL_SCAN_CHUNK, hi = lo + GLOBAL_SCAN_CHUNK < n ? lo + GLOBAL_SCAN_CHUNK : n;
  CurvePt acc = curve_infinity();
  for (size_t i = lo; i < hi; i++) acc = curve_add(acc, curve_load(pts + GLOBAL_POINT_WORDS * i), k);
  curve_store(totals + GLOBAL_POINT_WORDS * t, acc);
}
// scan, pass 2: chunk t of `pts` replaced by its inclusive sums, started from the scanned total of the chunks before it
KB_HD void global_chunk_rescan(u32* pts, size_t n, size_t t, const GlobalConsts& k, const u32* scanned_totals) {
  const size_t lo = t * GLOBAL_SCAN_CHUNK, hi = lo + GLOBAL_SCAN_CHUNK < n ? lo + GLOBAL_SCAN_CHUNK : n;
  CurvePt acc = t ? curve_load(scanned_totals + GLOBAL_POINT_WORDS * (t - 1)) : curve_infinity();
  for (size_t i = lo; i < hi; i++) {
    acc = curve_add(acc, curve_load(pts + GLOBAL_POINT_WORDS * i), k);
    curve_store(pts + GLOBAL_POINT_WORDS * i, acc);
  }
}

// finish: columns 64..98 of row `row` from the cumulative sums (sums[i] = start + the first i points, n + 1 of them), and for
// the rows past the last event the dummy lookup columns too (global/mod.rs:170-190, populate_real / populate_dummy)
KB_HD void global_finish_row(size_t row, size_t n, const u32* sums, const GlobalConsts& k, const GlobalOut& out) {
  if (row < n) {
    const u32* a = sums + GLOBAL_POINT_WORDS * row;
    for (int i = 0; i < 14; i++) { out(row, 64 + i, a[i]); out(row, 85 + i, a[14 + i]); }
    for (int i = 0; i < 7; i++) out(row, 78 + i, 0u);
    return;
  }
  // the final digest; with no event at all the reference's scan is empty and it is the dummy point (global/mod.rs:162-165)
  const CurvePt dummy = k.dummy, f = n ? curve_load(sums + GLOBAL_POINT_WORDS * n) : k.dummy;
  const Sep chk = curve_sum_checker_x(f, dummy, f);
  for (int i = 0; i < 16; i++) out(row, i, 0u);
  for (int i = 0; i < 7; i++) { out(row, 16 + i, dummy.x.c[i].v); out(row, 23 + i, dummy.y.c[i].v); }
  for (int i = 30; i < 64; i++) out(row, i, 0u);
  for (int i = 0; i < 7; i++) {
    out(row, 64 + i, f.x.c[i].v); out(row, 71 + i, f.y.c[i].v);
    out(row, 78 + i, chk.c[i].v);
    out(row, 85 + i, f.x.c[i].v); out(row, 92 + i, f.y.c[i].v);
  }
}

}  // namespace zkb

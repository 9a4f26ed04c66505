// KoalaBear (p = 2^31 - 2^24 + 1) in Montgomery form (R = 2^32) and its quartic extension
// F[x]/(x^4 - 3), for host and device.  The in-memory representation is the one the reference's
// RowMajorMatrix<KoalaBear> holds (constants cross-checked with
// crates/core/machine/include/kb31_t.hpp:27-34: MOD 0x7f000001, R mod p 0x1fffffe,
// R^2 mod p 0x17f7efe4; the reduction uses -p^-1 mod 2^32 = 0x7effffff, the negative of the
// header's MONTY_MU 0x81000001).  EF4 per crates/stark/src/air/extension.rs:55-75.
#pragma once
#if defined(__CUDACC_RTC__)
// NVRTC (run-time compiled constraint kernels, quotient_codegen.cpp): no host headers
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef unsigned long size_t;
#else
#include <cstdint>
#include <cstddef>
#endif

#if defined(__CUDACC__)
#define KB_HD __host__ __device__ __forceinline__
#else
#define KB_HD inline
#endif

namespace zkb {

typedef uint32_t u32;
typedef uint64_t u64;

constexpr u32 KB_P = 0x7f000001u;
constexpr u32 KB_ONE = 0x01fffffeu;    // R mod p
constexpr u32 KB_R2 = 0x17f7efe4u;     // R^2 mod p
constexpr u32 KB_GEN = 3;              // multiplicative generator (canonical)

struct Fp {
  u32 v;  // Montgomery residue in [0, p)
};

KB_HD Fp fp_raw(u32 v) { Fp r; r.v = v; return r; }
KB_HD Fp fp_zero() { return fp_raw(0); }
KB_HD Fp fp_one() { return fp_raw(KB_ONE); }

// 32 x 32 -> 64 product as one mul.wide (IMAD.WIDE): from C the compiler may widen an operand that
// came out of a 64-bit shift and emit a 64 x 64 product with stray zero-word adds
KB_HD u64 mul_wide(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
  u64 r;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
  return r;
#else
  return (u64)a * b;
#endif
}
// Montgomery reduction in the "plus" form: m = -x * p^-1 mod 2^32 makes x + m*p a multiple of 2^32,
// so the quotient is the high word of one multiply-accumulate (IMAD.HI with the 64-bit addend)
// and no separate subtraction is needed.  x < p * 2^32  =>  (x + m*p) / 2^32 in [0, 2p).
constexpr u32 KB_NPINV = 0x7effffffu;   // -p^-1 mod 2^32
KB_HD u32 mont_reduce_lazy(u64 x) {
#if defined(__CUDA_ARCH__)
  u32 m;   // kept a 32-bit multiply: left to the compiler it becomes a 64-bit product plus mask
  asm("mul.lo.u32 %0, %1, %2;" : "=r"(m) : "r"((u32)x), "r"(KB_NPINV));
#else
  u32 m = (u32)x * KB_NPINV;
#endif
  return (u32)((x + (u64)m * KB_P) >> 32);
}
KB_HD u32 mont_reduce(u64 x) {
  u32 r = mont_reduce_lazy(x);
  u32 t = r - KB_P;
  return t < r ? t : r;  // umin(r, r - p): one fused add-min on sm_100
}
// reduction of a sum of up to four products of residues (x < 4 p^2 < 2 p 2^32): one conditional
// subtraction of p * 2^32 on the high word, then the usual reduction
KB_HD u32 mont_reduce_wide(u64 x) {
  u32 hi = (u32)(x >> 32);
  u32 h2 = hi - KB_P;
  hi = h2 < hi ? h2 : hi;
  return mont_reduce(((u64)hi << 32) | (u32)x);
}
// a * b with a < 2p, b < p (product < p * 2^32)
KB_HD u32 mont_mul_raw(u32 a, u32 b) { return mont_reduce(mul_wide(a, b)); }
KB_HD Fp operator*(Fp a, Fp b) { return fp_raw(mont_reduce(mul_wide(a.v, b.v))); }
KB_HD Fp operator+(Fp a, Fp b) {
  u32 s = a.v + b.v;
  u32 t = s - KB_P;
  return fp_raw(t < s ? t : s);  // umin(s, s - p): s - p wraps high when s < p
}
KB_HD Fp operator-(Fp a, Fp b) {
  u32 s = a.v - b.v;
  u32 t = s + KB_P;
  return fp_raw(t < s ? t : s);  // umin(s, s + p)
}
KB_HD Fp operator-(Fp a) { return fp_raw(a.v ? KB_P - a.v : 0); }
KB_HD Fp& operator+=(Fp& a, Fp b) { a = a + b; return a; }
KB_HD Fp& operator-=(Fp& a, Fp b) { a = a - b; return a; }
KB_HD Fp& operator*=(Fp& a, Fp b) { a = a * b; return a; }
KB_HD bool operator==(Fp a, Fp b) { return a.v == b.v; }
KB_HD bool operator!=(Fp a, Fp b) { return a.v != b.v; }

KB_HD Fp fp_from_canonical(u32 x) { return fp_raw(mont_reduce((u64)x * KB_R2)); }
KB_HD u32 fp_to_canonical(Fp a) { return mont_reduce((u64)a.v); }
KB_HD Fp fp_double(Fp a) { return a + a; }

KB_HD Fp fp_pow(Fp b, u64 e) {
  Fp r = fp_one();
  while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
  return r;
}
KB_HD Fp fp_inv(Fp a) { return fp_pow(a, KB_P - 2); }
// halve: a/2
KB_HD Fp fp_halve(Fp a) { return fp_raw((a.v & 1) ? (u32)(((u64)a.v + KB_P) >> 1) : (a.v >> 1)); }

// ---- EF4 ---------------------------------------------------------------------------------
struct Ef {
  Fp c[4];
};
KB_HD Ef ef_zero() { Ef r; r.c[0] = r.c[1] = r.c[2] = r.c[3] = fp_zero(); return r; }
KB_HD Ef ef_from_fp(Fp a) { Ef r = ef_zero(); r.c[0] = a; return r; }
KB_HD Ef ef_one() { return ef_from_fp(fp_one()); }
KB_HD Ef operator+(const Ef& a, const Ef& b) { Ef r; for (int i = 0; i < 4; i++) r.c[i] = a.c[i] + b.c[i]; return r; }
KB_HD Ef operator-(const Ef& a, const Ef& b) { Ef r; for (int i = 0; i < 4; i++) r.c[i] = a.c[i] - b.c[i]; return r; }
KB_HD Ef operator-(const Ef& a) { Ef r; for (int i = 0; i < 4; i++) r.c[i] = -a.c[i]; return r; }
KB_HD Ef operator*(const Ef& a, Fp b) { Ef r; for (int i = 0; i < 4; i++) r.c[i] = a.c[i] * b; return r; }
KB_HD Ef operator+(const Ef& a, Fp b) { Ef r = a; r.c[0] += b; return r; }
KB_HD Ef operator-(const Ef& a, Fp b) { Ef r = a; r.c[0] -= b; return r; }
KB_HD Fp fp_mul3(Fp a) { return a + a + a; }
KB_HD Ef operator*(const Ef& a, const Ef& b) {
  Ef r;
  r.c[0] = a.c[0] * b.c[0] + fp_mul3(a.c[1] * b.c[3] + a.c[2] * b.c[2] + a.c[3] * b.c[1]);
  r.c[1] = a.c[0] * b.c[1] + a.c[1] * b.c[0] + fp_mul3(a.c[2] * b.c[3] + a.c[3] * b.c[2]);
  r.c[2] = a.c[0] * b.c[2] + a.c[1] * b.c[1] + a.c[2] * b.c[0] + fp_mul3(a.c[3] * b.c[3]);
  r.c[3] = a.c[0] * b.c[3] + a.c[1] * b.c[2] + a.c[2] * b.c[1] + a.c[3] * b.c[0];
  return r;
}
// Lazy EF accumulator: sums of (EF x base) products kept as raw 64-bit sums, reduced every
// fourth product.  acc.add(w, x) costs 4 multiply-accumulates instead of 4 modular multiply-adds.
struct EfAcc {
  u64 raw[4];
  Fp tot[4];
  int n;
  KB_HD void clear() { for (int i = 0; i < 4; i++) { raw[i] = 0; tot[i] = fp_raw(0); } n = 0; }
  KB_HD void flush() {
    for (int i = 0; i < 4; i++) { tot[i] = tot[i] + fp_raw(mont_reduce_wide(raw[i])); raw[i] = 0; }
    n = 0;
  }
  KB_HD void add(const Ef& w, Fp x) {
    for (int i = 0; i < 4; i++) raw[i] += (u64)w.c[i].v * x.v;
    if (++n == 4) flush();
  }
  KB_HD Ef value() { if (n) flush(); Ef r; for (int i = 0; i < 4; i++) r.c[i] = tot[i]; return r; }
};
KB_HD Ef& operator+=(Ef& a, const Ef& b) { a = a + b; return a; }
KB_HD Ef& operator-=(Ef& a, const Ef& b) { a = a - b; return a; }
KB_HD Ef& operator*=(Ef& a, const Ef& b) { a = a * b; return a; }
KB_HD bool ef_eq(const Ef& a, const Ef& b) { return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2] && a.c[3] == b.c[3]; }
KB_HD bool ef_is_zero(const Ef& a) { return (a.c[0].v | a.c[1].v | a.c[2].v | a.c[3].v) == 0; }

KB_HD Ef ef_pow(Ef b, u64 e) {
  Ef r = ef_one();
  while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
  return r;
}
// inverse via the norm down to F[y]/(y^2-3) (y = x^2) and then to F
KB_HD Ef ef_inv(const Ef& a) {
  Fp a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
  Fp A2_0 = a0 * a0 + fp_mul3(a2 * a2), A2_1 = fp_double(a0 * a2);
  Fp B2_0 = a1 * a1 + fp_mul3(a3 * a3), B2_1 = fp_double(a1 * a3);
  Fp N0 = A2_0 - fp_mul3(B2_1), N1 = A2_1 - B2_0;
  Fp d = fp_inv(N0 * N0 - fp_mul3(N1 * N1));
  Fp I0 = N0 * d, I1 = -(N1 * d);
  Ef r;
  r.c[0] = a0 * I0 + fp_mul3(a2 * I1);
  r.c[2] = a0 * I1 + a2 * I0;
  r.c[1] = -(a1 * I0 + fp_mul3(a3 * I1));
  r.c[3] = -(a1 * I1 + a3 * I0);
  return r;
}

// 1 / den for G denominators with ONE base-field inversion: ef_inv (kb31.cuh) takes the norm down to F and inverts
// there (a 30-step power), so the G norms share a Montgomery batch inversion.  A zero denominator (probability
// 2^-124 per lookup) has inverse 0 in ef_inv; here it is replaced by 1 in the product and masked afterwards, so the
// result is the same.
template <int G>
KB_HD void ef_inv_batch(const Ef* den, Ef* out) {
  Fp n0[G], n1[G], d[G], pref[G];
  for (int i = 0; i < G; i++) {
    const Fp a0 = den[i].c[0], a1 = den[i].c[1], a2 = den[i].c[2], a3 = den[i].c[3];
    const Fp A0 = a0 * a0 + fp_mul3(a2 * a2), A1 = fp_double(a0 * a2);
    const Fp B0 = a1 * a1 + fp_mul3(a3 * a3), B1 = fp_double(a1 * a3);
    n0[i] = A0 - fp_mul3(B1);
    n1[i] = A1 - B0;
    d[i] = n0[i] * n0[i] - fp_mul3(n1[i] * n1[i]);
    const Fp dd = d[i].v ? d[i] : fp_one();
    pref[i] = i ? pref[i - 1] * dd : dd;
  }
  Fp inv = fp_inv(pref[G - 1]);
  for (int i = G - 1; i >= 0; i--) {
    const Fp dd = d[i].v ? d[i] : fp_one();
    Fp di = i ? inv * pref[i - 1] : inv;
    inv = inv * dd;
    if (!d[i].v) di = fp_zero();
    const Fp I0 = n0[i] * di, I1 = -(n1[i] * di);
    const Fp a0 = den[i].c[0], a1 = den[i].c[1], a2 = den[i].c[2], a3 = den[i].c[3];
    out[i].c[0] = a0 * I0 + fp_mul3(a2 * I1);
    out[i].c[2] = a0 * I1 + a2 * I0;
    out[i].c[1] = -(a1 * I0 + fp_mul3(a3 * I1));
    out[i].c[3] = -(a1 * I1 + a3 * I0);
  }
}

KB_HD unsigned log2_exact(size_t n) { unsigned l = 0; while (((size_t)1 << l) < n) l++; return l; }
KB_HD u32 bitrev32(u32 x, unsigned bits) {
#if defined(__CUDA_ARCH__)
  return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
  u32 r = 0;
  for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
#endif
}

// element of multiplicative order 2^bits (Montgomery form); two_adic(24) = 3^127
KB_HD Fp two_adic_generator(unsigned bits) {
  Fp g = fp_pow(fp_from_canonical(KB_GEN), 127);
  for (unsigned i = bits; i < 24; i++) g *= g;
  return g;
}

}  // namespace zkb

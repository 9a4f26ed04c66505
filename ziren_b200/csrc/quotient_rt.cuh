// K3 device-side runtime shared by the bytecode interpreter (quotient.cu) and by the constraint kernels
// generated per chip and compiled with NVRTC when a context is created (quotient_codegen.cpp): kernel
// arguments, the row / selector prologue, the LogUp and global-sum constraints, the epilogue.
// Free of host headers (NVRTC compiles it from an embedded copy).
//
// One thread per row of the quotient domain GENERATOR * K_{n<<lqd}; thread t works on storage row t of the
// bit-reversed LDEs (natural index i = bitrev(t)), so the "local" loads of a warp are contiguous and the
// "next" row (i + 2^lqd) differs from t only in its top bits, i.e. is contiguous as well.
#pragma once
#include "kb31.cuh"
#include "machine_dev.h"

namespace zkb {

struct QuotArgs {
  const u32* prep; const u32* main_; const u32* perm;
  size_t H;
  u32 log_n, lqd;
  u32 ew, batch, main_width, global_scope;
  const Instr* code; u32 code_begin, code_end, n_air;
  const DevTerm* terms; const DevVPC* vpcs; const DevLookup* lookups; u32 lk_begin, lk_end;
  const DevFlatLookup* flk; const DevFlatTerm* fterms;      // the chip's own flattened lookups (machine_dev.h)
  const u32* lkK; const u32* lkE;                           // their per-proof coefficients (lookup_coefficients, logup.cu)
  const u32* alpha_pow;   // [C][4] Montgomery: alpha^(C-1-k)
  const u32* consts;      // constant pool (Montgomery)
  const u32* pub;
  const u32* tw_lo; const u32* tw_hi;
  Ef perm_alpha, local_sum;
  Ef bpow[17];
  u32 gsum[14];
  u32 zh[16], inv_zh[16];    // Z_H on the coset takes 2^lqd values
  u32 gen, ginv;             // GENERATOR, g_n^-1 (Montgomery)
  u32* out;
  // generated kernels of large chips run as `groups` CTAs per row tile (blockIdx.x = group, blockIdx.y = tile), each
  // evaluating a slice of the constraints; their partial sums ([group][4][Q] words) are combined by
  // quotient_combine_kernel.  groups <= 1: one CTA per tile (blockIdx.x), results written directly.
  u32 groups;
  u32* partial;
};

// K5 as generated code (quotient_codegen.cpp, kernel `lk`): the LogUp permutation trace rows of one chip.  prep / main_: column-major
// TRACES of n rows; out: n x 4E column-major (the batch columns; the running-sum column is filled by the scan kernels of
// logup.cu); rowsum: [4][n], the row's sum over its batches.
struct PermArgs {
  const u32* prep; const u32* main_;
  size_t n;
  const u32* lkK; const u32* lkE;
  u32* out; u32* rowsum;
};

struct QuotRow {
  size_t t, tn;            // storage rows of the local and the next row
  u32 i;                   // natural index on the quotient domain
  bool active;
  Fp is_first, is_last, is_trans;
};

__device__ __forceinline__ Fp q_tw_pow2(const u32* __restrict__ lo, const u32* __restrict__ hi, u32 E) {
  return fp_raw(__ldg(hi + (E >> 12))) * fp_raw(__ldg(lo + (E & 4095u)));
}
__device__ __forceinline__ Fp q_eval_vpc(const QuotArgs& a, u32 vi, size_t row) {
  DevVPC v = a.vpcs[vi];
  Fp acc = fp_raw(v.constant);
  for (u32 t = v.term_begin; t < v.term_end; t++) {
    DevTerm tm = a.terms[t];
    const u32* base = (tm.col & 0x80000000u) ? a.main_ : a.prep;
    acc += fp_raw(base[(size_t)(tm.col & 0x7fffffffu) * a.H + row]) * fp_raw(tm.w);
  }
  return acc;
}
__device__ __forceinline__ Ef load_ef(const u32* base, size_t H, u32 col4, size_t row) {
  Ef e;
#pragma unroll
  for (int c = 0; c < 4; c++) e.c[c] = fp_raw(base[(size_t)(col4 + c) * H + row]);
  return e;
}
__device__ __forceinline__ Ef load_apow(const u32* ap, u32 k) {
  uint4 v = __ldg(reinterpret_cast<const uint4*>(ap) + k);
  Ef e; e.c[0] = fp_raw(v.x); e.c[1] = fp_raw(v.y); e.c[2] = fp_raw(v.z); e.c[3] = fp_raw(v.w);
  return e;
}

// row indices and selectors at x = GENERATOR * w_Q^i   (crates/recursion/circuit/src/domain.rs:46-64)
__device__ __forceinline__ QuotRow q_prologue(const QuotArgs& a, size_t tile) {
  QuotRow r;
  const u32 lq = a.log_n + a.lqd;
  const size_t Q = (size_t)1 << lq;
  size_t t = tile * (size_t)blockDim.x + threadIdx.x;
  r.active = t < Q;
  if (!r.active) t = Q - 1;          // keep every thread in the barriers of the interpreter; result discarded
  r.t = t;
  r.i = bitrev32((u32)t, lq);
  r.tn = bitrev32((u32)((r.i + (1u << a.lqd)) & (Q - 1)), lq);
  Fp x = fp_raw(a.gen) * q_tw_pow2(a.tw_lo, a.tw_hi, r.i << (24 - lq));
  Fp zh = fp_raw(a.zh[r.i & ((1u << a.lqd) - 1)]);
  Fp d1 = x - fp_one(), d2 = x - fp_raw(a.ginv);
  Fp inv12 = fp_inv(d1 * d2);
  r.is_first = zh * (inv12 * d2); r.is_last = zh * (inv12 * d1); r.is_trans = d2;
  return r;
}

// fingerprint and signed multiplicity of the chip's lookup `lk` on row t: K + sum_t E_t * x[col_t] (machine_dev.h)
__device__ __forceinline__ void q_lookup(const QuotArgs& a, u32 lk, size_t t, Ef& rr, Fp& mult) {
  const DevFlatLookup l = a.flk[lk];
  rr = load_apow(a.lkK, lk);
  for (u32 ti = l.fterm_begin; ti < l.fterm_end; ti++) {
    const DevFlatTerm tm = a.fterms[ti];
    const u32* base = (tm.col & 0x80000000u) ? a.main_ : a.prep;
    rr += load_apow(a.lkE, ti) * fp_raw(__ldg(base + (size_t)(tm.col & 0x7fffffffu) * a.H + t));
  }
  const Fp mu = q_eval_vpc(a, l.mult_vpc, t);
  mult = l.is_send ? mu : -mu;
}

// permutation constraints (permutation.rs:205-347), then the global cumulative sum rows; k: index of the first
// LogUp constraint's alpha power.  The batches [b0, b1) are evaluated here: a kernel split into groups hands every
// group a range; the three running-sum constraints are linear in the batch entries, so every group adds its own
// entries' share and the group that owns the last batch (`tail`) adds the phi terms and the global-sum rows.
__device__ __forceinline__ void q_lookup_constraints(const QuotArgs& a, const QuotRow& r, Ef& acc, u32 k, u32 b0 = 0,
                                                     u32 b1 = 0xffffffffu, bool tail = true) {
  const size_t t = r.t, tn = r.tn;
  if (a.ew) {
    const u32 nlk = a.lk_end - a.lk_begin;
    const u32 k_sums = k + (a.ew - 1);
    if (b1 > a.ew - 1) b1 = a.ew - 1;
    u32 lk = b0 * a.batch;
    k += b0;
    Ef sum_local = ef_zero(), sum_next = ef_zero();
    for (u32 b = b0; b < b1; b++) {
      // product = prod_p rlc_p, numerator = sum_p mult_p prod_{q != p} rlc_q over the batch's lookups
      Ef product, numerator;
      if (a.batch == 1) {
        Fp m0;
        q_lookup(a, lk++, t, product, m0);
        numerator = ef_from_fp(m0);
      } else if (a.batch == 2) {
        Ef r0, r1 = ef_one();
        Fp m0, m1 = fp_zero();
        q_lookup(a, lk++, t, r0, m0);
        if (lk < nlk) q_lookup(a, lk++, t, r1, m1);
        product = r0 * r1;
        numerator = r1 * m0 + r0 * m1;
      } else {
        Ef rlc[8];
        Fp mult[8];
        u32 cnt = 0;
        for (; cnt < a.batch && lk < nlk; cnt++, lk++) q_lookup(a, lk, t, rlc[cnt], mult[cnt]);
        product = ef_one(); numerator = ef_zero();
        for (u32 p = 0; p < cnt; p++) {
          product *= rlc[p];
          Ef abc = ef_one();
          for (u32 q = 0; q < cnt; q++) if (q != p) abc *= rlc[q];
          numerator += abc * mult[p];
        }
      }
      Ef entry = load_ef(a.perm, a.H, 4 * b, t);
      acc += load_apow(a.alpha_pow, k++) * (product * entry - numerator);
      sum_local += entry;
      sum_next += load_ef(a.perm, a.H, 4 * b, tn);
    }
    Ef phi_local = ef_zero(), phi_next = ef_zero();
    if (tail) { phi_local = load_ef(a.perm, a.H, 4 * (a.ew - 1), t); phi_next = load_ef(a.perm, a.H, 4 * (a.ew - 1), tn); }
    k = k_sums;
    acc += load_apow(a.alpha_pow, k++) * ((phi_local - sum_local) * r.is_first);
    acc += load_apow(a.alpha_pow, k++) * ((phi_next - phi_local - sum_next) * r.is_trans);
    if (tail) acc += load_apow(a.alpha_pow, k) * ((phi_local - a.local_sum) * r.is_last);
    k++;
  }
  if (a.global_scope && tail) {
    for (int g = 0; g < 7; g++) {
      Fp mx = fp_raw(a.main_[(size_t)(a.main_width - 14 + g) * a.H + t]);
      Fp my = fp_raw(a.main_[(size_t)(a.main_width - 7 + g) * a.H + t]);
      acc += load_apow(a.alpha_pow, k++) * (r.is_last * (mx - fp_raw(a.gsum[g])));
      acc += load_apow(a.alpha_pow, k++) * (r.is_last * (my - fp_raw(a.gsum[7 + g])));
    }
  }
}

// What is left of the LogUp constraints once the batch constraints have been evaluated by GENERATED code
// (quotient_codegen.cpp, lookup shapes): the three running-sum constraints and the global cumulative sum rows.
// sum_local / sum_next: the sums of the batch entries on the local and the next row; k: alpha index of the first LogUp
// constraint.
struct QLookupSums { Ef acc, sum_local, sum_next; };
__device__ __forceinline__ void q_lookup_tail(const QuotArgs& a, const QuotRow& r, Ef& acc, u32 k, const Ef& sum_local,
                                              const Ef& sum_next) {
  const size_t t = r.t, tn = r.tn;
  if (a.ew) {
    k += a.ew - 1;
    const Ef phi_local = load_ef(a.perm, a.H, 4 * (a.ew - 1), t), phi_next = load_ef(a.perm, a.H, 4 * (a.ew - 1), tn);
    acc += load_apow(a.alpha_pow, k++) * ((phi_local - sum_local) * r.is_first);
    acc += load_apow(a.alpha_pow, k++) * ((phi_next - phi_local - sum_next) * r.is_trans);
    acc += load_apow(a.alpha_pow, k++) * ((phi_local - a.local_sum) * r.is_last);
  }
  if (a.global_scope) {
    for (int g = 0; g < 7; g++) {
      Fp mx = fp_raw(a.main_[(size_t)(a.main_width - 14 + g) * a.H + t]);
      Fp my = fp_raw(a.main_[(size_t)(a.main_width - 7 + g) * a.H + t]);
      acc += load_apow(a.alpha_pow, k++) * (r.is_last * (mx - fp_raw(a.gsum[g])));
      acc += load_apow(a.alpha_pow, k++) * (r.is_last * (my - fp_raw(a.gsum[7 + g])));
    }
  }
}

// quotient value and the split into 2^lqd chunks: chunk j = i mod 2^lqd, row k = i >> lqd
// (quotient_domain.split_evals, prover.rs:477-488)
__device__ __forceinline__ void q_epilogue(const QuotArgs& a, const QuotRow& r, const Ef& acc) {
  if (a.groups > 1) {          // this group's share of the row's sum; quotient_combine_kernel finishes the row
    const size_t Q = (size_t)1 << (a.log_n + a.lqd);
    if (r.active) {
#pragma unroll
      for (int c = 0; c < 4; c++) a.partial[((size_t)blockIdx.x * 4 + c) * Q + r.t] = acc.c[c].v;
    }
    return;
  }
  Ef q = acc * fp_raw(a.inv_zh[r.i & ((1u << a.lqd) - 1)]);
  const size_t n = (size_t)1 << a.log_n;
  u32* o = a.out + (size_t)(r.i & ((1u << a.lqd) - 1)) * 4 * n + (r.i >> a.lqd);
  if (r.active) {
#pragma unroll
    for (int c = 0; c < 4; c++) o[(size_t)c * n] = q.c[c].v;
  }
}

}  // namespace zkb

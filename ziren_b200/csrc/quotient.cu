// K3: constraint interpreter.  One thread per row of the quotient domain GENERATOR * K_{n<<lqd};
// thread t works on storage row t of the bit-reversed LDEs (natural index i = bitrev(t)), so the
// "local" loads of a warp are contiguous and the "next" row (i + 2^lqd) differs from t only in its
// top bits, i.e. is contiguous as well.  The chip's constraints arrive as register-allocated
// bytecode (machine.cpp); all threads execute the same instruction stream, so there is no
// divergence and instruction fetches are broadcast.
#include "quotient.h"

namespace zkb {

struct QuotArgs {
  const u32* prep; const u32* main_; const u32* perm;
  size_t H;
  u32 log_n, lqd;
  u32 ew, batch, main_width, global_scope;
  const Instr* code; u32 code_begin, code_end, n_air;
  const DevTerm* terms; const DevVPC* vpcs; const DevLookup* lookups; u32 lk_begin, lk_end;
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
};

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

constexpr int QCHUNK = 1024;   // 16 KB of bytecode per stage

template <int NREGS>
__global__ void __launch_bounds__(128) quotient_kernel(QuotArgs a) {
  const u32 lq = a.log_n + a.lqd;
  const size_t Q = (size_t)1 << lq;
  size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const bool active = t < Q;
  if (!active) t = Q - 1;          // keep every thread in the barriers below; result discarded
  const u32 i = bitrev32((u32)t, lq);
  const size_t tn = bitrev32((u32)((i + (1u << a.lqd)) & (Q - 1)), lq);

  // selectors at x = GENERATOR * w_Q^i   (crates/recursion/circuit/src/domain.rs:46-64)
  Fp x = fp_raw(a.gen) * tw_pow2(a.tw_lo, a.tw_hi, i << (24 - lq));
  Fp zh = fp_raw(a.zh[i & ((1u << a.lqd) - 1)]);
  Fp d1 = x - fp_one(), d2 = x - fp_raw(a.ginv);
  Fp inv12 = fp_inv(d1 * d2);
  Fp is_first = zh * (inv12 * d2), is_last = zh * (inv12 * d1), is_trans = d2;

  Fp regs[NREGS];
  EfAcc lazy;     // base-field constraints: alpha-power x value, reduced every fourth assert
  lazy.clear();
  // operands are tagged references (machine.h): registers, trace columns at the local / next
  // row, constants, public values, selectors
  auto fetch = [&](u32 o) -> Fp {
    const u32 idx = o & 0x1fffffffu;
    switch (o >> 29) {
      case O_REG: return regs[idx];
      case O_MAIN: return fp_raw(a.main_[(size_t)idx * a.H + t]);
      case O_MAIN_NEXT: return fp_raw(a.main_[(size_t)idx * a.H + tn]);
      case O_PREP: return fp_raw(a.prep[(size_t)idx * a.H + t]);
      case O_PREP_NEXT: return fp_raw(a.prep[(size_t)idx * a.H + tn]);
      case O_CONST: return fp_raw(__ldg(a.consts + idx));
      case O_PUB: return fp_raw(__ldg(a.pub + idx));
      default: return idx == 0 ? is_first : idx == 1 ? is_last : is_trans;
    }
  };
  // the bytecode is staged through shared memory in chunks: every thread runs the same
  // instruction stream, so one cooperative copy replaces a dependent global load per instruction
  __shared__ uint4 sh_code[QCHUNK];
  for (u32 cbase = a.code_begin; cbase < a.code_end; cbase += QCHUNK) {
    const u32 cn = min((u32)QCHUNK, a.code_end - cbase);
    __syncthreads();
    {
      const uint4* src = reinterpret_cast<const uint4*>(a.code + cbase);
      for (u32 w = threadIdx.x; w < cn; w += blockDim.x) sh_code[w] = __ldg(src + w);
    }
    __syncthreads();
    for (u32 pc = 0; pc < cn; pc++) {
      const uint4 ins = sh_code[pc];
      const u32 op = ins.x >> 24, dst = ins.x & 0xffffffu;
      if (op == I_ASSERT) { lazy.add(load_apow(a.alpha_pow, ins.w), fetch(ins.y)); continue; }
      if (op == I_ASSERT_SUB) { lazy.add(load_apow(a.alpha_pow, ins.w), fetch(ins.y) - fetch(ins.z)); continue; }
      const Fp x = fetch(ins.y);
      Fp v;
      if (op == I_NEG) v = -x;
      else {
        const Fp y = fetch(ins.z);
        v = op == I_ADD ? x + y : op == I_SUB ? x - y : x * y;
      }
      regs[dst] = v;
    }
  }

  Ef acc = lazy.value();
  // permutation constraints (permutation.rs:205-347)
  u32 k = a.n_air;
  if (a.ew) {
    u32 lk = a.lk_begin;
    Ef sum_local = ef_zero(), sum_next = ef_zero();
    for (u32 b = 0; b + 1 < a.ew; b++) {
      Ef rlc[8];
      Fp mult[8];
      u32 cnt = 0;
      for (; cnt < a.batch && lk < a.lk_end; cnt++, lk++) {
        DevLookup l = a.lookups[lk];
        Ef r = a.perm_alpha + fp_raw(l.kind);
        u32 j = 1;
        for (u32 vi = l.value_begin; vi < l.value_end; vi++, j++) r += a.bpow[j] * q_eval_vpc(a, vi, t);
        Fp mu = q_eval_vpc(a, l.mult_vpc, t);
        rlc[cnt] = r;
        mult[cnt] = l.is_send ? mu : -mu;
      }
      Ef product = ef_one(), numerator = ef_zero();
      for (u32 p = 0; p < cnt; p++) {
        product *= rlc[p];
        Ef abc = ef_one();
        for (u32 q = 0; q < cnt; q++) if (q != p) abc *= rlc[q];
        numerator += abc * mult[p];
      }
      Ef entry = load_ef(a.perm, a.H, 4 * b, t);
      acc += load_apow(a.alpha_pow, k++) * (product * entry - numerator);
      sum_local += entry;
      sum_next += load_ef(a.perm, a.H, 4 * b, tn);
    }
    Ef phi_local = load_ef(a.perm, a.H, 4 * (a.ew - 1), t), phi_next = load_ef(a.perm, a.H, 4 * (a.ew - 1), tn);
    acc += load_apow(a.alpha_pow, k++) * ((phi_local - sum_local) * is_first);
    acc += load_apow(a.alpha_pow, k++) * ((phi_next - phi_local - sum_next) * is_trans);
    acc += load_apow(a.alpha_pow, k++) * ((phi_local - a.local_sum) * is_last);
  }
  if (a.global_scope) {
    for (int g = 0; g < 7; g++) {
      Fp mx = fp_raw(a.main_[(size_t)(a.main_width - 14 + g) * a.H + t]);
      Fp my = fp_raw(a.main_[(size_t)(a.main_width - 7 + g) * a.H + t]);
      acc += load_apow(a.alpha_pow, k++) * (is_last * (mx - fp_raw(a.gsum[g])));
      acc += load_apow(a.alpha_pow, k++) * (is_last * (my - fp_raw(a.gsum[7 + g])));
    }
  }
  Ef q = acc * fp_raw(a.inv_zh[i & ((1u << a.lqd) - 1)]);
  // chunk j = i mod 2^lqd, row k = i >> lqd (quotient_domain.split_evals, prover.rs:477-488)
  const size_t n = (size_t)1 << a.log_n;
  u32* o = a.out + (size_t)(i & ((1u << a.lqd) - 1)) * 4 * n + (i >> a.lqd);
  if (active) {
#pragma unroll
    for (int c = 0; c < 4; c++) o[(size_t)c * n] = q.c[c].v;
  }
}

__global__ void alpha_pow_kernel(u32* out, u32 C, Ef alpha) {
  // out[k] = alpha^(C-1-k); C is at most a few thousand: one thread, sequential, runs once per chip
  if (blockIdx.x || threadIdx.x) return;
  Ef p = ef_one();
  for (u32 k = 0; k < C; k++) {
    u32* o = out + 4 * (size_t)(C - 1 - k);
    o[0] = p.c[0].v; o[1] = p.c[1].v; o[2] = p.c[2].v; o[3] = p.c[3].v;
    p *= alpha;
  }
}

void quotient_values(const MachineInfo& m, const ChipInfo& chip, const NttTables& tb, const QuotientInputs& in, u32* out,
                     cudaStream_t s) {
  QuotArgs a;
  a.prep = in.prep_lde; a.main_ = in.main_lde; a.perm = in.perm_lde; a.H = in.lde_h;
  a.log_n = in.log_n; a.lqd = chip.log_quotient_degree;
  a.ew = chip.perm_width_ef(); a.batch = chip.batch_size(); a.main_width = chip.main_width;
  a.global_scope = chip.global_scope ? 1 : 0;
  if (a.batch > 8) throw std::runtime_error("zkb200: LogUp batch size > 8 unsupported");
  if (a.lqd > 4) throw std::runtime_error("zkb200: log_quotient_degree > 4 unsupported");
  a.code = m.d_code; a.code_begin = chip.code_begin; a.code_end = chip.code_end; a.n_air = (u32)chip.constraints.size();
  a.terms = m.d_terms; a.vpcs = m.d_vpcs; a.lookups = m.d_lookups; a.lk_begin = chip.dev_lookup_begin; a.lk_end = chip.dev_lookup_end;
  a.pub = in.pub_dev; a.consts = m.d_consts; a.tw_lo = tb.tw_lo; a.tw_hi = tb.tw_hi;
  a.perm_alpha = in.perm_alpha; a.local_sum = in.local_sum;
  a.bpow[0] = ef_one();
  for (int i = 1; i < 17; i++) a.bpow[i] = a.bpow[i - 1] * in.perm_beta;
  memcpy(a.gsum, in.global_sum, sizeof(a.gsum));
  // Z_H(x) = x^n - 1 on x = g * w_Q^i:  g^n * (w_Q^n)^i, w_Q^n of order 2^lqd
  const size_t n = (size_t)1 << in.log_n;
  Fp gn = fp_pow(fp_from_canonical(KB_GEN), n);
  Fp wq = two_adic_generator(a.lqd);
  Fp cur = fp_one();
  for (u32 v = 0; v < (1u << a.lqd); v++) {
    Fp zh = gn * cur - fp_one();
    a.zh[v] = zh.v; a.inv_zh[v] = fp_inv(zh).v;
    cur *= wq;
  }
  a.gen = fp_from_canonical(KB_GEN).v;
  a.ginv = fp_inv(two_adic_generator(in.log_n)).v;
  const u32 C = chip.num_constraints();
  DevBuf apow((size_t)4 * (C ? C : 1), s);
  alpha_pow_kernel<<<1, 32, 0, s>>>(apow.p, C, in.alpha);
  ZKB_CHECK_LAUNCH();
  a.alpha_pow = apow.p;
  a.out = out;
  const size_t Q = n << a.lqd;
  const unsigned grid = ceil_div(Q, 128);
  if (chip.n_regs <= 32) quotient_kernel<32><<<grid, 128, 0, s>>>(a);
  else if (chip.n_regs <= 128) quotient_kernel<128><<<grid, 128, 0, s>>>(a);
  else if (chip.n_regs <= 512) quotient_kernel<512><<<grid, 128, 0, s>>>(a);
  else if (chip.n_regs <= 2048) quotient_kernel<2048><<<grid, 128, 0, s>>>(a);
  else throw std::runtime_error("zkb200: chip " + chip.name + " needs more than 2048 live constraint registers");
  ZKB_CHECK_LAUNCH();
}

}  // namespace zkb

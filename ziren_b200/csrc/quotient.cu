// K3: constraint interpreter.  One thread per row of the quotient domain GENERATOR * K_{n<<lqd};
// thread t works on storage row t of the bit-reversed LDEs (natural index i = bitrev(t)), so the
// "local" loads of a warp are contiguous and the "next" row (i + 2^lqd) differs from t only in its
// top bits, i.e. is contiguous as well.  The chip's constraints arrive as register-allocated
// bytecode (machine.cpp); all threads execute the same instruction stream, so there is no
// divergence and instruction fetches are broadcast.
#include "quotient.h"
#include "logup.h"
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include "quotient_codegen.h"
#include "quotient_rt.cuh"

namespace zkb {

std::atomic<unsigned long long> g_quotient_generated_launches{0}, g_quotient_interpreter_launches{0};
std::atomic<int> g_qk_block{256};            // zkb200_set_option("qk_block"): threads per CTA of the generated kernels
std::atomic<int> g_quotient_codegen{1};     // zkb200_set_option("quotient_codegen"); ZKB200_QUOTIENT=interp turns it off

constexpr int QCHUNK = 1024;   // 16 KB of bytecode per stage

template <int NREGS>
__global__ void __launch_bounds__(128) quotient_kernel(QuotArgs a) {
  const QuotRow row = q_prologue(a, blockIdx.x);
  const size_t t = row.t, tn = row.tn;
  const Fp is_first = row.is_first, is_last = row.is_last, is_trans = row.is_trans;

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
  q_lookup_constraints(a, row, acc, a.n_air);
  q_epilogue(a, row, acc);
}

__global__ void alpha_pow_kernel(u32* out, u32 C, Ef alpha) {
  // out[k] = alpha^(C-1-k), one thread per power (square and multiply): the real KeccakSponge chip has 3 970
  // constraints, a sequential chain of that many EF products took a millisecond per proof
  const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= C) return;
  const Ef p = ef_pow(alpha, C - 1 - k);
  u32* o = out + 4 * (size_t)k;
  o[0] = p.c[0].v; o[1] = p.c[1].v; o[2] = p.c[2].v; o[3] = p.c[3].v;
}

// Rows whose constraints were evaluated by `groups` CTAs each (generated kernels of large chips): sum the partial
// sums, divide by Z_H, split into chunks (q_epilogue).
__global__ void __launch_bounds__(256) quotient_combine_kernel(QuotArgs a) {
  const u32 lq = a.log_n + a.lqd;
  const size_t Q = (size_t)1 << lq;
  const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (t >= Q) return;
  Ef acc = ef_zero();
  for (u32 g = 0; g < a.groups; g++) {
    Ef p;
#pragma unroll
    for (int c = 0; c < 4; c++) p.c[c] = fp_raw(a.partial[((size_t)g * 4 + c) * Q + t]);
    acc += p;
  }
  QuotRow r;
  r.t = t; r.tn = t; r.active = true;
  r.i = bitrev32((u32)t, lq);
  r.is_first = r.is_last = r.is_trans = fp_zero();
  QuotArgs b = a;
  b.groups = 1;
  q_epilogue(b, r, acc);
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
  a.flk = m.d_flk + chip.dev_lookup_begin; a.fterms = m.d_fterms + chip.dev_fterm_begin;
  a.pub = in.pub_dev; a.consts = m.d_consts; a.tw_lo = tb.tw_lo; a.tw_hi = tb.tw_hi;
  a.perm_alpha = in.perm_alpha; a.local_sum = in.local_sum;
  a.bpow[0] = ef_one();
  for (int i = 1; i < 17; i++) a.bpow[i] = a.bpow[i - 1] * in.perm_beta;
  // per-proof fingerprint coefficients of the flattened lookups
  const u32 nlk = chip.dev_lookup_end - chip.dev_lookup_begin, nft = chip.dev_fterm_end - chip.dev_fterm_begin;
  DevBuf coefK(4 * (size_t)std::max<u32>(nlk, 1), s), coefE(4 * (size_t)std::max<u32>(nft, 1), s);
  if (a.ew) lookup_coefficients(m, chip, in.perm_alpha, a.bpow, coefK.p, coefE.p, s);
  a.lkK = coefK.p; a.lkE = coefE.p;
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
  alpha_pow_kernel<<<ceil_div(std::max<u32>(C, 1), 128), 128, 0, s>>>(apow.p, C, in.alpha);
  ZKB_CHECK_LAUNCH();
  a.alpha_pow = apow.p;
  a.out = out;
  a.groups = 1; a.partial = nullptr;
  const size_t Q = n << a.lqd;
  const unsigned grid = ceil_div(Q, 128);
  // generated straight-line kernel for this chip (NVRTC, quotient_codegen.cpp); the interpreter below is
  // the fallback and, with zkb200_set_option("quotient_codegen", 0), the implementation it is tested against
  if (g_quotient_codegen.load()) {
    if (void* k = quotient_generated_kernel(chip)) {
      void* params[] = {&a};
      const unsigned bs = (unsigned)std::max(32, std::min(256, g_qk_block.load()));
      // large chips: `groups` CTAs per row tile, each a slice of the constraints (quotient_codegen.cpp); the tile's
      // CTAs are neighbours in launch order, so the columns one slice loaded are in L2 when the next slice wants them
      const unsigned groups = quotient_codegen_groups(chip);
      DevBuf partial(groups > 1 ? (size_t)groups * 4 * Q : 0, s);
      a.groups = groups; a.partial = partial.p;
      const dim3 grid = groups > 1 ? dim3(groups, ceil_div(Q, bs)) : dim3(ceil_div(Q, bs));
      ZKB_CUDA(cudaLaunchKernel((const void*)k, grid, dim3(bs), params, 0, s));
      ZKB_CHECK_LAUNCH();
      if (groups > 1) {
        quotient_combine_kernel<<<ceil_div(Q, 256), 256, 0, s>>>(a);
        ZKB_CHECK_LAUNCH();
      }
      g_quotient_generated_launches++;
      return;
    }
  }
  g_quotient_interpreter_launches++;
  if (chip.n_regs <= 32) quotient_kernel<32><<<grid, 128, 0, s>>>(a);
  else if (chip.n_regs <= 128) quotient_kernel<128><<<grid, 128, 0, s>>>(a);
  else if (chip.n_regs <= 512) quotient_kernel<512><<<grid, 128, 0, s>>>(a);
  else if (chip.n_regs <= 2048) quotient_kernel<2048><<<grid, 128, 0, s>>>(a);
  else throw std::runtime_error("zkb200: chip " + chip.name + " needs more than 2048 live constraint registers");
  ZKB_CHECK_LAUNCH();
}

}  // namespace zkb

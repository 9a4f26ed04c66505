// K5: LogUp permutation trace.  One thread per row evaluates every lookup's fingerprint
// alpha + kind + sum_j beta^(j+1) v_j (VirtualPairCol linear combinations of the row), inverts it in
// EF4 and accumulates +-mult/fingerprint per batch; a three-phase modular prefix sum then fills the
// running-sum column.  Column-major traces keep every load of a warp contiguous.
#include "logup.h"

#include <atomic>
#include "quotient_codegen.h"
#include "quotient_rt.cuh"
namespace zkb {
std::atomic<int> g_logup_codegen{1};         // zkb200_set_option("logup_codegen"): K5 from the generated module (1) or the run-time loop (0)

struct LogupArgs {
  const u32* prep; const u32* main_; size_t n;
  const DevTerm* terms; const DevVPC* vpcs;
  const DevFlatLookup* flk; const DevFlatTerm* fterms;      // the chip's own ranges
  const u32* lkK; const u32* lkE;                           // per-proof coefficients (lookup_coefficients)
  u32 nlk, batch, ew;
  u32* out;         // n x 4*ew column-major
  u32* rowsum;      // 4 x n
};

__device__ __forceinline__ Ef load_ef4(const u32* p, u32 i) {
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + i);
  Ef e; e.c[0] = fp_raw(v.x); e.c[1] = fp_raw(v.y); e.c[2] = fp_raw(v.z); e.c[3] = fp_raw(v.w);
  return e;
}
__device__ __forceinline__ Fp eval_vpc(const LogupArgs& a, u32 vi, size_t r) {
  DevVPC v = a.vpcs[vi];
  Fp acc = fp_raw(v.constant);
  for (u32 t = v.term_begin; t < v.term_end; t++) {
    DevTerm tm = a.terms[t];
    const u32* base = (tm.col & 0x80000000u) ? a.main_ : a.prep;
    Fp x = fp_raw(base[(size_t)(tm.col & 0x7fffffffu) * a.n + r]);
    acc += x * fp_raw(tm.w);
  }
  return acc;
}

__global__ void __launch_bounds__(128) logup_rows_kernel(LogupArgs a) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= a.n) return;
  constexpr int G = 8;                       // lookups per shared inversion; the batch size (1, 2 or 4) divides it
  Ef total = ef_zero();
  u32 b = 0;                                 // permutation column being filled
  for (u32 lk0 = 0; lk0 < a.nlk; lk0 += G) {
    Ef den[G], inv[G];
    Fp mult[G];
#pragma unroll
    for (int k = 0; k < G; k++) {
      const u32 lk = lk0 + k;
      if (lk < a.nlk) {
        const DevFlatLookup l = a.flk[lk];
        Ef dk = load_ef4(a.lkK, lk);
        for (u32 t = l.fterm_begin; t < l.fterm_end; t++) {
          const DevFlatTerm tm = a.fterms[t];
          const u32* base = (tm.col & 0x80000000u) ? a.main_ : a.prep;
          dk += load_ef4(a.lkE, t) * fp_raw(base[(size_t)(tm.col & 0x7fffffffu) * a.n + r]);
        }
        Fp m = eval_vpc(a, l.mult_vpc, r);
        mult[k] = l.is_send ? m : -m;
        den[k] = dk;
      } else { den[k] = ef_one(); mult[k] = fp_zero(); }
    }
    ef_inv_batch<G>(den, inv);
    Ef v = ef_zero();
#pragma unroll
    for (int k = 0; k < G; k++) {
      if (lk0 + k < a.nlk) {
        v += inv[k] * mult[k];
        // a batch is complete after `batch` lookups (a power of two dividing G), or at the chip's last lookup
        if ((((u32)k + 1) & (a.batch - 1)) == 0 || lk0 + k + 1 == a.nlk) {
#pragma unroll
          for (int c = 0; c < 4; c++) a.out[(size_t)(4 * b + c) * a.n + r] = v.c[c].v;
          total += v;
          v = ef_zero();
          b++;
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; c++) a.rowsum[(size_t)c * a.n + r] = total.c[c].v;
}

// ---- per-proof fingerprint coefficients of the flattened lookups (machine_dev.h) ----------------------------
struct CoefArgs {
  const DevLookup* lookups; const DevVPC* vpcs; const DevFlatTerm* fterms;
  u32 nlk, nterms;
  Ef alpha;
  Ef bpow[17];
  u32* K; u32* E;
};
__global__ void lookup_coefficients_kernel(CoefArgs a) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.nlk) {
    const DevLookup l = a.lookups[i];
    Ef k = a.alpha + fp_raw(l.kind);
    u32 j = 1;
    for (u32 vi = l.value_begin; vi < l.value_end; vi++, j++) k += a.bpow[j] * fp_raw(a.vpcs[vi].constant);
#pragma unroll
    for (int c = 0; c < 4; c++) a.K[4 * i + c] = k.c[c].v;
  }
  if (i < a.nterms) {
    const DevFlatTerm t = a.fterms[i];
    const Ef e = a.bpow[t.j] * fp_raw(t.w);
#pragma unroll
    for (int c = 0; c < 4; c++) a.E[4 * i + c] = e.c[c].v;
  }
}
void lookup_coefficients(const MachineInfo& m, const ChipInfo& chip, const Ef& alpha, const Ef* bpow17, u32* K_out, u32* E_out,
                         cudaStream_t s) {
  CoefArgs a;
  a.lookups = m.d_lookups + chip.dev_lookup_begin; a.vpcs = m.d_vpcs; a.fterms = m.d_fterms + chip.dev_fterm_begin;
  a.nlk = chip.dev_lookup_end - chip.dev_lookup_begin; a.nterms = chip.dev_fterm_end - chip.dev_fterm_begin;
  a.alpha = alpha;
  for (int i = 0; i < 17; i++) a.bpow[i] = bpow17[i];
  a.K = K_out; a.E = E_out;
  const u32 n = std::max(a.nlk, a.nterms);
  if (!n) return;
  lookup_coefficients_kernel<<<ceil_div(n, 128), 128, 0, s>>>(a);
  ZKB_CHECK_LAUNCH();
}

// ---- modular inclusive prefix sum over 4 independent sequences of length n ----------------------
constexpr int SCAN_BLOCK = 1024;
__device__ __forceinline__ Fp block_scan_incl(Fp v, Fp* sh) {
  // Hillis-Steele in shared memory (1024 threads)
  int t = threadIdx.x;
  sh[t] = v;
  __syncthreads();
  for (int d = 1; d < SCAN_BLOCK; d <<= 1) {
    Fp x = sh[t];
    if (t >= d) x = x + sh[t - d];
    __syncthreads();
    sh[t] = x;
    __syncthreads();
  }
  return sh[t];
}
__global__ void __launch_bounds__(SCAN_BLOCK) scan_phase1(const u32* in, u32* out, u32* block_tot, size_t n, size_t nblk) {
  __shared__ Fp sh[SCAN_BLOCK];
  size_t seq = blockIdx.y, i = blockIdx.x * (size_t)SCAN_BLOCK + threadIdx.x;
  Fp v = i < n ? fp_raw(in[seq * n + i]) : fp_zero();
  Fp r = block_scan_incl(v, sh);
  if (i < n) out[seq * n + i] = r.v;
  if (threadIdx.x == SCAN_BLOCK - 1) block_tot[seq * nblk + blockIdx.x] = r.v;
}
__global__ void __launch_bounds__(SCAN_BLOCK) scan_phase2(u32* block_tot, size_t nblk) {
  // exclusive scan of the block totals of one sequence, one CTA per sequence
  __shared__ Fp sh[SCAN_BLOCK];
  __shared__ Fp carry;
  u32* bt = block_tot + blockIdx.x * nblk;
  if (threadIdx.x == 0) carry = fp_zero();
  __syncthreads();
  for (size_t base = 0; base < nblk; base += SCAN_BLOCK) {
    size_t i = base + threadIdx.x;
    Fp v = i < nblk ? fp_raw(bt[i]) : fp_zero();
    Fp incl = block_scan_incl(v, sh);
    Fp c = carry;
    if (i < nblk) bt[i] = (c + incl - v).v;
    __syncthreads();
    if (threadIdx.x == SCAN_BLOCK - 1) carry = c + incl;
    __syncthreads();
  }
}
__global__ void __launch_bounds__(SCAN_BLOCK) scan_phase3(u32* out, const u32* block_tot, size_t n, size_t nblk, u32* total_out) {
  size_t seq = blockIdx.y, i = blockIdx.x * (size_t)SCAN_BLOCK + threadIdx.x;
  if (i >= n) return;
  Fp r = fp_raw(out[seq * n + i]) + fp_raw(block_tot[seq * nblk + blockIdx.x]);
  out[seq * n + i] = r.v;
  if (i == n - 1) total_out[seq] = r.v;
}

void permutation_trace(const MachineInfo& m, const ChipInfo& chip, const u32* prep, const u32* main_, size_t n,
                       const Ef& alpha, const Ef& beta, u32* out, u32* local_sum_dev, cudaStream_t s) {
  const u32 ew = chip.perm_width_ef();
  if (ew == 0) { ZKB_CUDA(cudaMemsetAsync(local_sum_dev, 0, 16, s)); return; }
  if (chip.max_values > 16) throw std::runtime_error("zkb200: lookup tuple longer than 16 values");
  if (chip.batch_size() > 8) throw std::runtime_error("zkb200: permutation batches of more than 8 lookups (log_quotient_degree > 3)");
  LogupArgs a;
  a.prep = prep; a.main_ = main_; a.n = n;
  a.terms = m.d_terms; a.vpcs = m.d_vpcs;
  a.flk = m.d_flk + chip.dev_lookup_begin; a.fterms = m.d_fterms + chip.dev_fterm_begin;
  a.nlk = chip.dev_lookup_end - chip.dev_lookup_begin; a.batch = chip.batch_size(); a.ew = ew;
  Ef bpow[17];
  bpow[0] = ef_one();
  for (int i = 1; i < 17; i++) bpow[i] = bpow[i - 1] * beta;
  const u32 nterms = chip.dev_fterm_end - chip.dev_fterm_begin;
  DevBuf coefK(4 * (size_t)std::max<u32>(a.nlk, 1), s), coefE(4 * (size_t)std::max<u32>(nterms, 1), s);
  lookup_coefficients(m, chip, alpha, bpow, coefK.p, coefE.p, s);
  a.lkK = coefK.p; a.lkE = coefE.p;
  a.out = out;
  const size_t nblk = ceil_div(n, SCAN_BLOCK);
  DevBuf rowsum(4 * n, s), btot(4 * nblk, s);
  a.rowsum = rowsum.p;
  // the chip's generated K5 kernel (quotient_codegen.cpp: one loop per batch shape, loads first), else the data-driven loop
  void* gen = g_logup_codegen.load() ? permutation_generated_kernel(chip) : nullptr;
  if (gen) {
    PermArgs pa;
    pa.prep = prep; pa.main_ = main_; pa.n = n; pa.lkK = coefK.p; pa.lkE = coefE.p; pa.out = out; pa.rowsum = rowsum.p;
    void* params[] = {&pa};
    ZKB_CUDA(cudaLaunchKernel((const void*)gen, dim3(ceil_div(n, 128)), dim3(128), params, 0, s));
  } else {
    logup_rows_kernel<<<ceil_div(n, 128), 128, 0, s>>>(a);
  }
  ZKB_CHECK_LAUNCH();
  u32* last = out + (size_t)4 * (ew - 1) * n;   // the 4 running-sum columns are contiguous
  scan_phase1<<<dim3((unsigned)nblk, 4), SCAN_BLOCK, 0, s>>>(rowsum.p, last, btot.p, n, nblk);
  ZKB_CHECK_LAUNCH();
  scan_phase2<<<4, SCAN_BLOCK, 0, s>>>(btot.p, nblk);
  ZKB_CHECK_LAUNCH();
  scan_phase3<<<dim3((unsigned)nblk, 4), SCAN_BLOCK, 0, s>>>(last, btot.p, n, nblk, local_sum_dev);
  ZKB_CHECK_LAUNCH();
}

}  // namespace zkb

// K5: LogUp permutation trace.  One thread per row evaluates every lookup's fingerprint
// alpha + kind + sum_j beta^(j+1) v_j (VirtualPairCol linear combinations of the row), inverts it in
// EF4 and accumulates +-mult/fingerprint per batch; a three-phase modular prefix sum then fills the
// running-sum column.  Column-major traces keep every load of a warp contiguous.
#include "logup.h"

namespace zkb {

struct LogupArgs {
  const u32* prep; const u32* main_; size_t n;
  const DevTerm* terms; const DevVPC* vpcs; const DevLookup* lookups;
  u32 lk_begin, lk_end, batch, ew;
  Ef alpha;
  Ef bpow[17];      // beta^0 .. beta^16
  u32* out;         // n x 4*ew column-major
  u32* rowsum;      // 4 x n
};

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
  Ef total = ef_zero();
  u32 lk = a.lk_begin;
  for (u32 b = 0; b + 1 < a.ew; b++) {
    Ef v = ef_zero();
    for (u32 k = 0; k < a.batch && lk < a.lk_end; k++, lk++) {
      DevLookup l = a.lookups[lk];
      Ef den = a.alpha + fp_raw(l.kind);
      u32 j = 1;
      for (u32 vi = l.value_begin; vi < l.value_end; vi++, j++) den += a.bpow[j] * eval_vpc(a, vi, r);
      Fp mult = eval_vpc(a, l.mult_vpc, r);
      if (!l.is_send) mult = -mult;
      v += ef_inv(den) * mult;
    }
#pragma unroll
    for (int c = 0; c < 4; c++) a.out[(size_t)(4 * b + c) * a.n + r] = v.c[c].v;
    total += v;
  }
#pragma unroll
  for (int c = 0; c < 4; c++) a.rowsum[(size_t)c * a.n + r] = total.c[c].v;
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
  LogupArgs a;
  a.prep = prep; a.main_ = main_; a.n = n;
  a.terms = m.d_terms; a.vpcs = m.d_vpcs; a.lookups = m.d_lookups;
  a.lk_begin = chip.dev_lookup_begin; a.lk_end = chip.dev_lookup_end; a.batch = chip.batch_size(); a.ew = ew;
  a.alpha = alpha;
  a.bpow[0] = ef_one();
  for (int i = 1; i < 17; i++) a.bpow[i] = a.bpow[i - 1] * beta;
  a.out = out;
  const size_t nblk = ceil_div(n, SCAN_BLOCK);
  DevBuf rowsum(4 * n, s), btot(4 * nblk, s);
  a.rowsum = rowsum.p;
  logup_rows_kernel<<<ceil_div(n, 128), 128, 0, s>>>(a);
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

#include "fri.h"

namespace zkb {

__global__ void __launch_bounds__(128) fri_fold_kernel(const u32* tw_lo, const u32* tw_hi, const u32* __restrict__ in, size_t m,
                                                       unsigned log_m, Ef beta, Ef beta2, const u32* __restrict__ beta_dev, Fp neg_half, Fp half,
                                                       const u32* __restrict__ ro_next, u32* __restrict__ out) {
  const size_t hm = m >> 1;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= hm) return;
  if (beta_dev) {            // beta sampled by the device-resident challenger (hash.cu)
#pragma unroll
    for (int c = 0; c < 4; c++) beta.c[c] = fp_raw(__ldg(beta_dev + c));
    beta2 = beta * beta;
  }
  Ef e0, e1;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    uint2 pr = reinterpret_cast<const uint2*>(in + (size_t)c * m)[i];
    e0.c[c] = fp_raw(pr.x);
    e1.c[c] = fp_raw(pr.y);
  }
  // x0 = w_m^bitrev(2i); 1/(x1 - x0) = -1/(2 x0) = -1/2 * w_m^(-bitrev(2i))
  u32 e = bitrev32((u32)(2 * i), log_m);
  u32 E = (0u - (e << (24 - log_m))) & ((1u << 24) - 1);
  Fp cinv = neg_half * tw_pow2(tw_lo, tw_hi, E);
  // (beta - x0) * c = beta * c + 1/2
  Ef f = beta * cinv + half;
  Ef r = e0 + (e1 - e0) * f;
  if (ro_next) {
    Ef q;
#pragma unroll
    for (int c = 0; c < 4; c++) q.c[c] = fp_raw(ro_next[(size_t)c * hm + i]);
    r += beta2 * q;
  }
#pragma unroll
  for (int c = 0; c < 4; c++) out[(size_t)c * hm + i] = r.c[c].v;
}
void fri_fold(const NttTables& tb, const u32* in, size_t m, const Ef& beta, const u32* ro_next, u32* out, cudaStream_t s) {
  const size_t hm = m >> 1;
  Fp half = fp_halve(fp_one());
  fri_fold_kernel<<<ceil_div(hm, 128), 128, 0, s>>>(tb.tw_lo, tb.tw_hi, in, m, log2_exact(m), beta, beta * beta, nullptr, -half, half,
                                                    ro_next, out);
  ZKB_CHECK_LAUNCH();
}
void fri_fold_dev_beta(const NttTables& tb, const u32* in, size_t m, const u32* beta_dev, const u32* ro_next, u32* out, cudaStream_t s) {
  const size_t hm = m >> 1;
  Fp half = fp_halve(fp_one());
  fri_fold_kernel<<<ceil_div(hm, 128), 128, 0, s>>>(tb.tw_lo, tb.tw_hi, in, m, log2_exact(m), ef_zero(), ef_zero(), beta_dev, -half, half,
                                                    ro_next, out);
  ZKB_CHECK_LAUNCH();
}

__global__ void __launch_bounds__(256) gather_kernel(const GatherJob* __restrict__ jobs, size_t njobs, u32* __restrict__ dst) {
  size_t j = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  if (j >= njobs) return;
  GatherJob job = jobs[j];
  for (u32 k = threadIdx.x & 31; k < job.count; k += 32) dst[job.dst + k] = fp_to_canonical(fp_raw(job.src[(size_t)k * job.stride]));
}
void gather_canonical(const GatherJob* jobs_dev, size_t njobs, u32* dst, cudaStream_t s) {
  if (!njobs) return;
  gather_kernel<<<ceil_div(njobs * 32, 256), 256, 0, s>>>(jobs_dev, njobs, dst);
  ZKB_CHECK_LAUNCH();
}

}  // namespace zkb

// Host build of a GENERATED constraint kernel (K3, ziren_b200/csrc/quotient_codegen.cpp): the CUDA source the generator
// writes for one chip is compiled here as plain C++ behind a few shims (no device qualifiers, __ldg = load, blockIdx /
// threadIdx as variables) and run row by row, so that the CPU suite can compare what the generator emitted - constraint
// shapes, parameter tables, LogUp batch shapes, the run-time header quotient_rt.cuh - with the oracle's quotient without a
// GPU.  Test infrastructure only (built by tests/test_codegen_host.py with -DQK_SOURCE="<dumped .cu>").
#include <cstddef>
#include <cstdint>
#include <cstring>
#define __device__
#define __global__
#define __noinline__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
struct QkDim { unsigned x = 0, y = 0, z = 0; };
static QkDim blockIdx, threadIdx, blockDim, gridDim;
template <class T> static inline T __ldg(const T* p) { return *p; }

#include QK_SOURCE

using namespace zkb;

extern "C" {
// All field elements Montgomery.  LDEs column-major with H rows (bit-reversed), out: 2^lqd chunk matrices of n x 4 column-major.
int qk_host_run(unsigned log_n, unsigned lqd, size_t H, const uint32_t* prep, const uint32_t* main_, const uint32_t* perm,
                unsigned ew, unsigned batch, unsigned main_width, unsigned global_scope, unsigned n_air, unsigned n_lookups,
                const uint32_t* alpha_pow, const uint32_t* lkK, const uint32_t* lkE, const uint32_t* pub,
                const uint32_t* tw_lo, const uint32_t* tw_hi, const uint32_t* local_sum, const uint32_t* gsum,
                const uint32_t* zh, const uint32_t* inv_zh, uint32_t gen, uint32_t ginv, uint32_t* out) {
  QuotArgs a;
  memset(&a, 0, sizeof(a));
  a.prep = prep; a.main_ = main_; a.perm = perm; a.H = H;
  a.log_n = log_n; a.lqd = lqd; a.ew = ew; a.batch = batch; a.main_width = main_width; a.global_scope = global_scope;
  a.n_air = n_air; a.lk_begin = 0; a.lk_end = n_lookups;
  a.alpha_pow = alpha_pow; a.lkK = lkK; a.lkE = lkE; a.pub = pub; a.tw_lo = tw_lo; a.tw_hi = tw_hi;
  for (int i = 0; i < 4; i++) a.local_sum.c[i] = fp_raw(local_sum[i]);
  memcpy(a.gsum, gsum, sizeof(a.gsum));
  for (unsigned v = 0; v < (1u << lqd); v++) { a.zh[v] = zh[v]; a.inv_zh[v] = inv_zh[v]; }
  a.gen = gen; a.ginv = ginv; a.out = out;
  a.groups = 1; a.partial = nullptr;
  const size_t Q = (size_t)1 << (log_n + lqd);
  blockDim.x = 256;
  for (size_t b = 0; b < (Q + 255) / 256; b++)
    for (unsigned t = 0; t < 256; t++) {
      blockIdx.x = (unsigned)b; threadIdx.x = t;
      qk(a);
    }
  return 0;
}
#ifdef QK_HAS_LK
// the generated K5 kernel `lk` of the same module: traces column-major with n rows; out: n x 4E column-major, rowsum [4][n]
int lk_host_run(size_t n, const uint32_t* prep, const uint32_t* main_, const uint32_t* lkK, const uint32_t* lkE, uint32_t* out,
                uint32_t* rowsum) {
  PermArgs a;
  a.prep = prep; a.main_ = main_; a.n = n; a.lkK = lkK; a.lkE = lkE; a.out = out; a.rowsum = rowsum;
  blockDim.x = 128;
  for (size_t b = 0; b < (n + 127) / 128; b++)
    for (unsigned t = 0; t < 128; t++) {
      blockIdx.x = (unsigned)b; threadIdx.x = t;
      lk(a);
    }
  return 0;
}
#endif
}

// K1: batched KoalaBear NTT / coset low-degree extension over column-major matrices.
// Replaces [P3-upstream] Radix2DitParallel::coset_lde_batch + bit_reverse_rows as called from
// TwoAdicFriPcs::commit (reference call sites crates/stark/src/prover.rs:277,403,497 and
// crates/stark/src/machine.rs:416-417).
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include <vector>
#include "common.h"

namespace zkb {

struct NttTables {
  u32* tw_lo = nullptr;  // w^e, e in [0, 4096)           (w = generator of the 2^24 subgroup)
  u32* tw_hi = nullptr;  // w^(4096 e), e in [0, 4096)
  void* small_tw = nullptr;                       // Shoup pairs of w_{2^K}^(+-e), K <= 12
  mutable std::map<int, void*> four_step;         // per (log n, direction): four-step twiddles by position
  // per (log n, blow-up, shift): coset scale factors by position.  Entries are reference counted: a
  // caller parks the pointer it got in its lane's keep-alive list until that lane has drained, so
  // evicting the cache can never free a table that queued kernels still read.
  mutable std::map<u64, std::shared_ptr<void>> scale_cache;
  mutable std::map<u64, size_t> scale_bytes;
  mutable size_t scale_cache_bytes = 0;
  mutable std::mutex mu;                          // tables are shared by the compute lanes
  void init(cudaStream_t s);
  void destroy();
  const void* four_step_table(int K1, int logS, bool inverse, cudaStream_t s) const;
  std::shared_ptr<void> scale_table(unsigned log_n, unsigned log_blowup, Fp shift, cudaStream_t s) const;
};

// Coset LDE of every column: in = evaluations over H_n in natural order (col-major n x w,
// column stride in_stride), out = evaluations over shift * K_{n << log_blowup} stored
// BIT-REVERSED by row (col-major, column stride out_stride).  Montgomery residues.
// `keep`: tables the queued kernels read; the caller drops them once the stream has drained.
typedef std::vector<std::shared_ptr<void>> KeepAlive;
void coset_lde_batch(const NttTables& tb, const u32* in, size_t in_stride, u32* out, size_t out_stride,
                     unsigned log_n, size_t width, unsigned log_blowup, Fp shift, cudaStream_t s, KeepAlive& keep);

// Plain DFT of every column, natural order in; out natural (bitrev_out = false) or bit-reversed.
void ntt_batch(const NttTables& tb, const u32* in, u32* out, unsigned log_n, size_t width, bool inverse,
               bool bitrev_out, cudaStream_t s);

#if defined(__CUDACC__)
// w^E for E in [0, 2^24) from the two-level table (w generates the 2^24 subgroup)
__device__ __forceinline__ Fp tw_pow2(const u32* __restrict__ lo, const u32* __restrict__ hi, u32 E) {
  return fp_raw(__ldg(hi + (E >> 12))) * fp_raw(__ldg(lo + (E & 4095u)));
}
#endif

// element-wise helpers used by the PCS
void bitrev_rows(const u32* in, u32* out, unsigned log_n, size_t width, cudaStream_t s);

}  // namespace zkb

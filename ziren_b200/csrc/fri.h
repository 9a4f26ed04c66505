// K4c/K4e: FRI fold and the query-phase gathers ([P3-upstream] fri::prover::{commit_phase,
// answer_query}; pinned by crates/recursion/circuit/src/fri.rs:247-361 and :107-128).
#pragma once
#include "common.h"
#include "ntt.h"

namespace zkb {
// out[i] = e0 + (beta - x0)(e1 - e0)/(x1 - x0) (+ beta^2 * ro_next[i]);  in: [4][m], out: [4][m/2]
void fri_fold(const NttTables& tb, const u32* in, size_t m, const Ef& beta, const u32* ro_next, u32* out, cudaStream_t s);
// the same with beta read from DEVICE memory (4 Montgomery words written by the device-resident challenger)
void fri_fold_dev_beta(const NttTables& tb, const u32* in, size_t m, const u32* beta_dev, const u32* ro_next, u32* out, cudaStream_t s);

// dst[job.dst + k] = canonical(src[k * stride]), k < count
struct GatherJob { const u32* src; u64 stride; u32 count; u32 dst; };
void gather_canonical(const GatherJob* jobs_dev, size_t njobs, u32* dst, cudaStream_t s);
}  // namespace zkb

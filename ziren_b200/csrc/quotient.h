// K3: per-row constraint / quotient evaluation on the quotient coset (quotient_values,
// crates/stark/src/quotient.rs:19-171 with ProverConstraintFolder crates/stark/src/folder.rs:52-149
// and eval_permutation_constraints crates/stark/src/permutation.rs:205-347).  The reference
// materialises three row-major copies of the LDEs per chip (prover.rs:435-445); here the committed
// column-major, bit-reversed LDEs are read in place.
#pragma once
#include "common.h"
#include "machine.h"
#include "ntt.h"

namespace zkb {
struct QuotientInputs {
  const u32* prep_lde;   // column-major, height lde_h, bit-reversed rows; null if the chip has none
  const u32* main_lde;
  const u32* perm_lde;
  size_t lde_h;
  unsigned log_n;
  Ef perm_alpha, perm_beta, local_sum, alpha;
  u32 global_sum[14];    // Montgomery
  const u32* pub_dev;    // Montgomery public values on the device
};
// out: 2^lqd chunk matrices, each column-major n x 4 (natural row order), chunk j at out + j*4n
void quotient_values(const MachineInfo& m, const ChipInfo& chip, const NttTables& tb, const QuotientInputs& in, u32* out,
                     cudaStream_t s);
}  // namespace zkb

// K5: LogUp permutation trace on the device (generate_permutation_trace,
// crates/stark/src/permutation.rs:102-196, driven from crates/stark/src/prover.rs:337-365).
#pragma once
#include "common.h"
#include "machine.h"

namespace zkb {
// prep/main: column-major traces of height n (prep may be null).  out: column-major n x 4E
// (EF flattened to base, last EF column = running sum).  local_sum_dev[4] receives the total.
void permutation_trace(const MachineInfo& m, const ChipInfo& chip, const u32* prep, const u32* main_, size_t n,
                       const Ef& alpha, const Ef& beta, u32* out, u32* local_sum_dev, cudaStream_t s);
// Per proof and chip: the fingerprint coefficients of the flattened lookups (machine_dev.h): K_out[nlk][4] and
// E_out[nterms][4] (Montgomery), for alpha and the powers of beta.
void lookup_coefficients(const MachineInfo& m, const ChipInfo& chip, const Ef& alpha, const Ef* bpow17, u32* K_out, u32* E_out,
                         cudaStream_t s);
}  // namespace zkb

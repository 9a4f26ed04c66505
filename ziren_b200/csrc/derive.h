// K7: multiplicity columns of a receive-only table from the rows of its senders (derive.cuh).
#pragma once
#include "machine.h"
#include "derive.cuh"

namespace zkb {

struct DeriveSender { const ChipInfo* chip; DeriveTable table; };
// out: receiver_height x receiver.main_width words, column-major Montgomery (device); returns the number of lookups counted.
// Throws if the receiver has no receive made of preprocessed columns with a plain main column as its multiplicity, or if a
// sender's tuple of one of the receiver's kinds is in none of its rows.
u64 derive_multiplicities(const MachineInfo& m, const ChipInfo& receiver, const u32* receiver_prep, size_t receiver_height,
                          const std::vector<DeriveSender>& senders, u32* out, cudaStream_t s);

}  // namespace zkb

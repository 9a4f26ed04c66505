// K3, generated code: one constraint kernel per chip, written out as straight-line CUDA from the chip's
// symbolic constraint DAG and compiled for sm_100a with NVRTC the first time the chip is proved
// (cubins are cached on disk).  The reference monomorphises `chip.eval(&mut folder)` in Rust
// (crates/stark/src/quotient.rs:157, folder.rs:52-149); across the C ABI the constraints are data, so
// the specialisation happens here, at run time.  The bytecode interpreter (quotient.cu) stays as the
// fallback for a host without libnvrtc and as the second implementation the parity tests compare with.
#pragma once
#include <string>
#include "machine.h"

namespace zkb {

// CUDA source of the chip's kernel (extern "C" __global__ void qk(QuotArgs)); pure function of the chip
std::string quotient_kernel_source(const ChipInfo& chip);

// Compiled kernel handle (cudaKernel_t as void*) for the chip, or nullptr when run-time compilation is
// unavailable / failed (the reason is kept in quotient_codegen_last_error()).  Thread safe; results are
// memoised per source text for the life of the process.
void* quotient_generated_kernel(const ChipInfo& chip);
// `lk` of the same module: the chip's LogUp permutation-trace rows (K5), or nullptr (no lookups / compilation unavailable)
void* permutation_generated_kernel(const ChipInfo& chip);
const char* quotient_codegen_last_error();
// How many CTAs share a row tile in the chip's generated kernel (1: the whole constraint set in one CTA).  Pure
// function of the chip, the same number the generator wrote the kernel for.
unsigned quotient_codegen_groups(const ChipInfo& chip);
// NVRTC only (no device needed): size of the chip's sm_100a cubin, 0 on failure.  Used by the CPU tests.
size_t quotient_codegen_compile_only(const ChipInfo& chip);

}  // namespace zkb

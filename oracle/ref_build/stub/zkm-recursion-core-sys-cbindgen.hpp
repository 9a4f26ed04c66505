// Stand-in for the header cbindgen generates in the reference's Rust build
// (crates/recursion/core/build.rs); that build cannot run here (no Rust toolchain).  It only
// declares the layout types and size constants the reference's Poseidon2 headers name.  This
// file is ours; the reference's headers are compiled from where they lie under /root/reference.
#pragma once
#include <cstddef>
#include <cstdint>
#include "kb31_t.hpp"

namespace zkm_recursion_core_sys {
constexpr size_t WIDTH = 16;
constexpr size_t NUM_EXTERNAL_ROUNDS = 8;
constexpr size_t NUM_INTERNAL_ROUNDS = 13;
constexpr size_t NUM_INTERNAL_ROUNDS_S0 = NUM_INTERNAL_ROUNDS - 1;
constexpr size_t OUTPUT_ROUND_IDX = NUM_EXTERNAL_ROUNDS + 2;  // src/chips/poseidon2_skinny/trace.rs:47

template <class F> struct Poseidon2Event { F input[WIDTH]; F output[WIDTH]; };
template <class F> struct Poseidon2 { F state_var[WIDTH]; F internal_rounds_s0[NUM_INTERNAL_ROUNDS_S0]; };
template <class F> struct Poseidon2Io { F input[WIDTH]; F output[WIDTH]; };
template <class F> struct Poseidon2Instr { Poseidon2Io<F> addrs; F mults[WIDTH]; };
template <class F> struct MemoryPreprocessed { F addr; F mult; };
template <class F> struct RoundCountersPreprocessed {
  F is_input_round, is_external_round, is_internal_round; F round_constants[WIDTH];
};
template <class F> struct Poseidon2PreprocessedColsSkinny {
  MemoryPreprocessed<F> memory_preprocessed[WIDTH];
  RoundCountersPreprocessed<F> round_counters_preprocessed;
};
}  // namespace zkm_recursion_core_sys

// ORACLE SUPPORT (test infrastructure).  Thin C shim over the reference's OWN C++ headers,
// compiled from /root/reference where they lie (never copied into this repo):
//   crates/core/machine/include/kb31_t.hpp                  (host branch, :451-623)
//   crates/recursion/core/include/poseidon2_skinny.hpp      (event_to_row, :51-75)
//   crates/recursion/core/include/poseidon2.hpp, poseidon2_constants.hpp
// Output goes to oracle/_ref/libzkref.so; tests use it to check the restated field arithmetic
// and Poseidon2 permutation against the reference itself.
#include "kb31_t.hpp"
#include "poseidon2_skinny.hpp"

using namespace zkm_recursion_core_sys;

extern "C" {
uint32_t ref_kb31_mul(uint32_t a, uint32_t b) {
  return (kb31_t::from_canonical_u32(a) * kb31_t::from_canonical_u32(b)).as_canonical_u32();
}
uint32_t ref_kb31_add(uint32_t a, uint32_t b) {
  return (kb31_t::from_canonical_u32(a) + kb31_t::from_canonical_u32(b)).as_canonical_u32();
}
uint32_t ref_kb31_sub(uint32_t a, uint32_t b) {
  return (kb31_t::from_canonical_u32(a) - kb31_t::from_canonical_u32(b)).as_canonical_u32();
}
uint32_t ref_kb31_inv(uint32_t a) { return kb31_t::from_canonical_u32(a).reciprocal().as_canonical_u32(); }
uint32_t ref_kb31_to_monty(uint32_t a) { return kb31_t::to_monty(a); }
uint32_t ref_kb31_from_monty(uint32_t a) { return kb31_t::from_monty(a); }
// the reference's trace generator runs the whole permutation; the last row holds the output
void ref_poseidon2_permute(uint32_t* state) {
  Poseidon2Event<kb31_t> ev;
  for (int i = 0; i < 16; i++) ev.input[i] = kb31_t::from_canonical_u32(state[i]);
  Poseidon2<kb31_t> cols[OUTPUT_ROUND_IDX + 1];
  poseidon2_skinny::event_to_row<kb31_t>(ev, cols);
  for (int i = 0; i < 16; i++) state[i] = cols[OUTPUT_ROUND_IDX].state_var[i].as_canonical_u32();
}
}

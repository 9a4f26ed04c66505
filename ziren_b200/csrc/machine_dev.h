// Device images of the machine tables (see machine.h).  Kept free of host headers: this file is also
// compiled by NVRTC for the run-time generated constraint kernels (quotient_codegen.cpp).
#pragma once
#include "kb31.cuh"

namespace zkb {

// ---- device-side tables -----------------------------------------------------------------------
struct DevTerm { u32 col; u32 w; };                 // col: bit 31 set = main trace, else preprocessed; w Montgomery
struct DevVPC { u32 constant; u32 term_begin, term_end; };
struct DevLookup { u32 kind; u32 is_send; u32 mult_vpc; u32 value_begin, value_end; };   // kind Montgomery

// Flattened lookups: the fingerprint of a lookup is affine in the row,
//   alpha + kind + sum_j beta^j v_j(row) = K + sum_t E_t * x[col_t],   K = alpha + kind + sum_j beta^j const_j,  E_t = beta^j w_t,
// so K and the E_t are computed ONCE per proof (lookup_coefficients, logup.cu) and the row kernels (K5 permutation
// trace, K3 LogUp constraints) run one flat loop of EF x base multiply-adds per lookup instead of walking
// DevLookup -> DevVPC -> DevTerm for every row.  fterm_* index the chip's own term range.
struct DevFlatLookup { u32 fterm_begin, fterm_end; u32 mult_vpc; u32 is_send; };
struct DevFlatTerm { u32 col; u32 j; u32 w; };      // col as DevTerm; j: 1-based position of the value in the tuple; w Montgomery

// Bytecode of the constraint interpreter (K3).  16-byte instructions {op|dst, a, b, c}; operands
// are tagged references, so trace columns, public values, selectors and constants are read where
// they are used instead of through separate load instructions:
//   operand = kind << 29 | index     kind 0 register, 1 main local col, 2 main next col,
//                                         3 prep local col, 4 prep next col, 5 constant-pool slot,
//                                         6 public value, 7 selector (0 first, 1 last, 2 transition)
//   op 0 ADD  1 SUB  2 MUL : r[dst] = a (op) b          3 NEG : r[dst] = -a
//   op 4 ASSERT            : acc += alpha_pow[c] * a     (assert_zero, folder.rs:79-84)
//   op 5 ASSERT_SUB        : acc += alpha_pow[c] * (a - b)   (assert_eq / "x*y - z" in one step)
struct Instr { u32 op_dst; u32 a, b, c; };
enum InstrOp : u32 { I_ADD = 0, I_SUB = 1, I_MUL = 2, I_NEG = 3, I_ASSERT = 4, I_ASSERT_SUB = 5 };
enum OperandKind : u32 { O_REG = 0, O_MAIN = 1, O_MAIN_NEXT = 2, O_PREP = 3, O_PREP_NEXT = 4, O_CONST = 5, O_PUB = 6, O_SEL = 7 };

}  // namespace zkb

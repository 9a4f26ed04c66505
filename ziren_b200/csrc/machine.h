// Static description of the machine (chips as data) on the host and its device image.
// Parses the "ZKMD" descriptor documented in include/zkb200.h and lowers every chip to what the
// kernels consume: lookup tables for the LogUp kernels and a register-allocated bytecode for the
// constraint interpreter (K3).
//
// Reference counterparts: Chip<F, A> (crates/stark/src/chip.rs:19-176), Lookup / VirtualPairCol
// (crates/stark/src/lookup/lookup.rs:10-19), the symbolic constraint DAG of
// get_symbolic_constraints (crates/stark/src/machine.rs:377-389).
#pragma once
#include "common.h"
#include <string>
#include <vector>

namespace zkb {

enum NodeOp : u32 { N_CONST = 0, N_MAIN, N_PREP, N_PUB, N_IS_FIRST, N_IS_LAST, N_IS_TRANS, N_ADD, N_SUB, N_MUL, N_NEG };

struct HostTerm { bool is_main; u32 col; u32 w_canon; };
struct HostVPC { u32 const_canon; std::vector<HostTerm> terms; };
struct HostLookup { u32 kind, scope; HostVPC mult; std::vector<HostVPC> values; bool is_send; };
struct HostNode { u32 op, a, b; };

// ---- device-side tables -----------------------------------------------------------------------
struct DevTerm { u32 col; u32 w; };                 // col: bit 31 set = main trace, else preprocessed; w Montgomery
struct DevVPC { u32 constant; u32 term_begin, term_end; };
struct DevLookup { u32 kind; u32 is_send; u32 mult_vpc; u32 value_begin, value_end; };   // kind Montgomery

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

struct ChipInfo {
  std::string name;
  u32 prep_width = 0, main_width = 0, log_quotient_degree = 1;
  bool local_only = false, global_scope = false;
  std::vector<HostLookup> lookups;       // local-scope lookups in evaluation order: sends, then receives
  std::vector<HostNode> nodes;
  std::vector<u32> constraints;
  u32 n_sends_total = 0, n_receives_total = 0;

  u32 batch_size() const { return 1u << log_quotient_degree; }
  u32 perm_width_ef() const { u32 n = (u32)lookups.size(); return n ? (n + batch_size() - 1) / batch_size() + 1 : 0; }
  u32 num_constraints() const {
    u32 c = (u32)constraints.size();
    if (perm_width_ef()) c += perm_width_ef() - 1 + 3;
    if (global_scope) c += 14;
    return c;
  }

  // device image (filled by MachineInfo::upload)
  u32 dev_lookup_begin = 0, dev_lookup_end = 0;   // range in the machine-wide DevLookup table
  u32 max_values = 0;                             // longest lookup tuple
  u32 code_begin = 0, code_end = 0;               // range in the machine-wide Instr table
  u32 n_regs = 0;
  u32 const_begin = 0;                            // first slot of this chip in the constant pool
};

struct MachineInfo {
  std::vector<ChipInfo> chips;
  u32 num_pv_elts = 0, log_blowup = 1, num_queries = 84, pow_bits = 16;

  // device tables
  DevTerm* d_terms = nullptr;
  DevVPC* d_vpcs = nullptr;
  DevLookup* d_lookups = nullptr;
  Instr* d_code = nullptr;
  u32* d_consts = nullptr;      // constant pool, Montgomery

  void parse(const u32* words, size_t n);
  void upload();       // lower + copy tables to the current device
  void destroy();
  const ChipInfo* find(const std::string& name) const {
    for (auto& c : chips) if (c.name == name) return &c;
    return nullptr;
  }
};

}  // namespace zkb

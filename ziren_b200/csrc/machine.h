// Static description of the machine (chips as data) on the host and its device image.
// Parses the "ZKMD" descriptor documented in include/zkb200.h and lowers every chip to what the
// kernels consume: lookup tables for the LogUp kernels and a register-allocated bytecode for the
// constraint interpreter (K3).
//
// Reference counterparts: Chip<F, A> (crates/stark/src/chip.rs:19-176), Lookup / VirtualPairCol
// (crates/stark/src/lookup/lookup.rs:10-19), the symbolic constraint DAG of
// get_symbolic_constraints (crates/stark/src/machine.rs:377-389).
#pragma once
#include <atomic>
#include "common.h"
#include "machine_dev.h"
#include <string>
#include <vector>

namespace zkb {

enum NodeOp : u32 { N_CONST = 0, N_MAIN, N_PREP, N_PUB, N_IS_FIRST, N_IS_LAST, N_IS_TRANS, N_ADD, N_SUB, N_MUL, N_NEG };

struct HostTerm { bool is_main; u32 col; u32 w_canon; };
struct HostVPC { u32 const_canon; std::vector<HostTerm> terms; };
struct HostLookup { u32 kind, scope; HostVPC mult; std::vector<HostVPC> values; bool is_send; };
struct HostNode { u32 op, a, b; };

// a per-chip cache slot that copies as "empty" (ChipInfo stays copyable)
struct KernelSlot {
  mutable std::atomic<void*> p{nullptr};
  mutable std::atomic<void*> lk{nullptr};       // the module's K5 kernel, same encoding as p
  mutable std::atomic<unsigned> groups{0};      // CTAs per row tile of the generated kernel, 0 = not computed yet
  KernelSlot() = default;
  KernelSlot(const KernelSlot&) {}
  KernelSlot& operator=(const KernelSlot&) { p.store(nullptr); lk.store(nullptr); groups.store(0); return *this; }
};

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
  u32 dev_lookup_begin = 0, dev_lookup_end = 0;   // range in the machine-wide DevLookup / DevFlatLookup tables
  u32 dev_fterm_begin = 0, dev_fterm_end = 0;     // range in the machine-wide DevFlatTerm table
  u32 max_values = 0;                             // longest lookup tuple
  u32 code_begin = 0, code_end = 0;               // range in the machine-wide Instr table
  u32 n_regs = 0;
  u32 const_begin = 0;                            // first slot of this chip in the constant pool
  // generated constraint kernel of this chip (quotient_codegen.cpp), looked up once: null = not yet, 1 = none (interpreter)
  KernelSlot qk_cached;
};

struct MachineInfo {
  std::vector<ChipInfo> chips;
  u32 num_pv_elts = 0, log_blowup = 1, num_queries = 84, pow_bits = 16;

  // device tables
  DevTerm* d_terms = nullptr;
  DevVPC* d_vpcs = nullptr;
  DevLookup* d_lookups = nullptr;
  DevFlatLookup* d_flk = nullptr;       // per lookup, parallel to d_lookups
  DevFlatTerm* d_fterms = nullptr;
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

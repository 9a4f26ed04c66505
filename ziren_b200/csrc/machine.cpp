#include "machine.h"
#include <algorithm>

namespace zkb {

namespace {
struct Reader {
  const u32* p; size_t n, pos = 0;
  u32 next() { if (pos >= n) throw std::runtime_error("zkb200: machine descriptor truncated"); return p[pos++]; }
  std::string str() {
    u32 len = next();
    std::string s;
    for (u32 i = 0; i < (len + 3) / 4; i++) {
      u32 w = next();
      for (int b = 0; b < 4; b++) if (s.size() < len) s.push_back((char)((w >> (8 * b)) & 255));
    }
    return s;
  }
  HostVPC vpc() {
    HostVPC v;
    v.const_canon = next();
    u32 nt = next();
    for (u32 i = 0; i < nt; i++) { HostTerm t; t.is_main = next() != 0; t.col = next(); t.w_canon = next(); v.terms.push_back(t); }
    return v;
  }
};
}  // namespace

void MachineInfo::parse(const u32* words, size_t n) {
  Reader r{words, n};
  if (r.next() != 0x444d4b5au) throw std::runtime_error("zkb200: bad machine descriptor magic");
  if (r.next() != 1) throw std::runtime_error("zkb200: unsupported machine descriptor version");
  u32 nchips = r.next();
  num_pv_elts = r.next();
  log_blowup = r.next();
  num_queries = r.next();
  pow_bits = r.next();
  if (log_blowup < 1 || log_blowup > 4) throw std::runtime_error("zkb200: log_blowup out of range");
  for (u32 ci = 0; ci < nchips; ci++) {
    ChipInfo c;
    c.name = r.str();
    c.prep_width = r.next(); c.main_width = r.next(); c.log_quotient_degree = r.next();
    c.local_only = r.next() != 0; c.global_scope = r.next() != 0;
    u32 ns = r.next(), nr = r.next(), nn = r.next(), nc = r.next();
    c.n_sends_total = ns; c.n_receives_total = nr;
    for (u32 i = 0; i < ns + nr; i++) {
      HostLookup l;
      l.kind = r.next(); l.scope = r.next();
      u32 nv = r.next();
      l.mult = r.vpc();
      for (u32 j = 0; j < nv; j++) l.values.push_back(r.vpc());
      l.is_send = i < ns;
      auto check = [&](const HostVPC& v) {
        for (auto& t : v.terms)
          if (t.col >= (t.is_main ? c.main_width : c.prep_width)) throw std::runtime_error("zkb200: lookup column out of range in chip " + c.name);
      };
      check(l.mult);
      for (auto& v : l.values) check(v);
      // bpow[] of the LogUp and quotient kernels holds beta^0..beta^16
      if (l.values.size() > 16) throw std::runtime_error("zkb200: lookup tuple longer than 16 values in chip " + c.name);
      if (l.scope == 0) c.lookups.push_back(std::move(l));   // only local lookups enter the permutation
    }
    for (u32 i = 0; i < nn; i++) {
      HostNode nd; nd.op = r.next(); nd.a = r.next(); nd.b = r.next();
      if (nd.op > N_NEG) throw std::runtime_error("zkb200: bad node opcode in chip " + c.name);
      if (nd.op >= N_ADD && (nd.a >= i || (nd.op != N_NEG && nd.b >= i))) throw std::runtime_error("zkb200: constraint DAG is not topologically ordered");
      if (nd.op == N_MAIN && nd.a >= c.main_width) throw std::runtime_error("zkb200: main column out of range in chip " + c.name);
      if (nd.op == N_PUB && nd.a >= num_pv_elts) throw std::runtime_error("zkb200: public value index out of range in chip " + c.name);
      if (nd.op == N_PREP && nd.a >= c.prep_width) throw std::runtime_error("zkb200: preprocessed column out of range in chip " + c.name);
      c.nodes.push_back(nd);
    }
    for (u32 i = 0; i < nc; i++) { u32 x = r.next(); if (x >= nn) throw std::runtime_error("zkb200: constraint id out of range"); c.constraints.push_back(x); }
    if (c.log_quotient_degree > log_blowup) throw std::runtime_error("zkb200: chip " + c.name + " needs log_quotient_degree > log_blowup");
    if (c.global_scope && c.main_width < 14) throw std::runtime_error("zkb200: global-scope chip narrower than 14 columns");
    chips.push_back(std::move(c));
  }
}

// Lower one chip's constraint DAG to interpreter bytecode: leaves become tagged operands, inner
// nodes get registers by linear scan, a SUB that is only asserted becomes ASSERT_SUB.
static void lower_chip(ChipInfo& c, std::vector<Instr>& code, std::vector<u32>& consts) {
  const size_t nn = c.nodes.size();
  std::vector<char> reach(nn, 0);
  for (u32 k : c.constraints) reach[k] = 1;
  for (size_t i = nn; i-- > 0;) {
    if (!reach[i]) continue;
    const HostNode& nd = c.nodes[i];
    if (nd.op >= N_ADD) { reach[nd.a] = 1; if (nd.op != N_NEG) reach[nd.b] = 1; }
  }
  std::vector<u32> uses(nn, 0);   // remaining consumers among op nodes
  for (size_t i = 0; i < nn; i++) {
    if (!reach[i]) continue;
    const HostNode& nd = c.nodes[i];
    if (nd.op >= N_ADD) { uses[nd.a]++; if (nd.op != N_NEG) uses[nd.b]++; }
  }
  std::vector<std::vector<u32>> asserts_of(nn);
  for (u32 k = 0; k < c.constraints.size(); k++) asserts_of[c.constraints[k]].push_back(k);

  std::vector<int> reg(nn, -1);
  std::vector<u32> free_regs;
  u32 n_regs = 0;
  c.const_begin = (u32)consts.size();
  auto alloc = [&]() { if (!free_regs.empty()) { u32 r = free_regs.back(); free_regs.pop_back(); return r; } return n_regs++; };
  auto emit = [&](u32 op, u32 dst, u32 a, u32 b, u32 k) { code.push_back(Instr{(op << 24) | dst, a, b, k}); };
  auto tag = [](u32 kind, u32 idx) {
    if (idx >= (1u << 29)) throw std::runtime_error("zkb200: operand index out of range");
    return (kind << 29) | idx;
  };
  auto operand = [&](u32 id) -> u32 {
    const HostNode& nd = c.nodes[id];
    switch (nd.op) {
      case N_CONST: consts.push_back(fp_from_canonical(nd.a % KB_P).v); return tag(O_CONST, (u32)consts.size() - 1);
      case N_MAIN: return tag(nd.b ? O_MAIN_NEXT : O_MAIN, nd.a);
      case N_PREP: return tag(nd.b ? O_PREP_NEXT : O_PREP, nd.a);
      case N_PUB: return tag(O_PUB, nd.a);
      case N_IS_FIRST: return tag(O_SEL, 0);
      case N_IS_LAST: return tag(O_SEL, 1);
      case N_IS_TRANS: return tag(O_SEL, 2);
      default:
        if (reg[id] < 0) throw std::runtime_error("zkb200: internal: operand used before definition");
        return tag(O_REG, (u32)reg[id]);
    }
  };
  auto release = [&](u32 id) {
    if (c.nodes[id].op < N_ADD) return;          // leaves hold no register
    if (--uses[id] == 0) free_regs.push_back((u32)reg[id]);
  };
  c.code_begin = (u32)code.size();
  for (size_t i = 0; i < nn; i++) {
    if (!reach[i]) continue;
    const HostNode& nd = c.nodes[i];
    if (nd.op < N_ADD) {
      for (u32 k : asserts_of[i]) emit(I_ASSERT, 0, operand((u32)i), 0, k);
      continue;
    }
    const u32 oa = operand(nd.a), ob = nd.op != N_NEG ? operand(nd.b) : 0;
    if (nd.op == N_SUB && uses[i] == 0) {
      // only asserted: fold the subtraction into the assert
      for (u32 k : asserts_of[i]) emit(I_ASSERT_SUB, 0, oa, ob, k);
      release(nd.a); release(nd.b);
      continue;
    }
    release(nd.a);
    if (nd.op != N_NEG) release(nd.b);
    const u32 r = alloc();
    reg[i] = (int)r;
    const u32 op = nd.op == N_ADD ? I_ADD : nd.op == N_SUB ? I_SUB : nd.op == N_MUL ? I_MUL : I_NEG;
    emit(op, r, oa, ob, 0);
    for (u32 k : asserts_of[i]) emit(I_ASSERT, 0, tag(O_REG, r), 0, k);
    if (uses[i] == 0) free_regs.push_back(r);
  }
  c.code_end = (u32)code.size();
  c.n_regs = n_regs ? n_regs : 1;
}

void MachineInfo::upload() {
  std::vector<DevTerm> terms;
  std::vector<DevVPC> vpcs;
  std::vector<DevLookup> lookups;
  std::vector<Instr> code;
  std::vector<u32> consts;
  std::vector<DevFlatLookup> flk;
  std::vector<DevFlatTerm> fterms;
  auto add_vpc = [&](const HostVPC& v) {
    DevVPC d;
    d.constant = fp_from_canonical(v.const_canon % KB_P).v;
    d.term_begin = (u32)terms.size();
    for (auto& t : v.terms) terms.push_back(DevTerm{t.col | (t.is_main ? 0x80000000u : 0u), fp_from_canonical(t.w_canon % KB_P).v});
    d.term_end = (u32)terms.size();
    vpcs.push_back(d);
    return (u32)vpcs.size() - 1;
  };
  for (auto& c : chips) {
    c.dev_lookup_begin = (u32)lookups.size();
    c.dev_fterm_begin = (u32)fterms.size();
    c.max_values = 0;
    for (auto& l : c.lookups) {
      DevFlatLookup f;
      f.fterm_begin = (u32)fterms.size() - c.dev_fterm_begin;
      for (size_t j = 0; j < l.values.size(); j++)
        for (auto& t : l.values[j].terms)
          fterms.push_back(DevFlatTerm{t.col | (t.is_main ? 0x80000000u : 0u), (u32)j + 1, fp_from_canonical(t.w_canon % KB_P).v});
      f.fterm_end = (u32)fterms.size() - c.dev_fterm_begin;
      f.is_send = l.is_send ? 1 : 0;
      DevLookup d;
      d.kind = fp_from_canonical(l.kind).v;
      d.is_send = l.is_send ? 1 : 0;
      d.mult_vpc = add_vpc(l.mult);
      d.value_begin = (u32)vpcs.size();
      for (auto& v : l.values) add_vpc(v);
      d.value_end = (u32)vpcs.size();
      c.max_values = std::max<u32>(c.max_values, (u32)l.values.size());
      lookups.push_back(d);
      f.mult_vpc = d.mult_vpc;
      flk.push_back(f);
    }
    c.dev_lookup_end = (u32)lookups.size();
    c.dev_fterm_end = (u32)fterms.size();
    lower_chip(c, code, consts);
  }
  auto up = [](auto*& dptr, const auto& v) {
    using T = typename std::remove_reference<decltype(v)>::type::value_type;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    ZKB_CUDA(cudaMalloc((void**)&dptr, bytes));
    if (!v.empty()) ZKB_CUDA(cudaMemcpy(dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  };
  up(d_terms, terms);
  up(d_vpcs, vpcs);
  up(d_lookups, lookups);
  up(d_flk, flk);
  up(d_fterms, fterms);
  up(d_code, code);
  up(d_consts, consts);
}

void MachineInfo::destroy() {
  if (d_terms) cudaFree(d_terms);
  if (d_vpcs) cudaFree(d_vpcs);
  if (d_lookups) cudaFree(d_lookups);
  if (d_flk) cudaFree(d_flk);
  if (d_fterms) cudaFree(d_fterms);
  d_flk = nullptr; d_fterms = nullptr;
  if (d_code) cudaFree(d_code);
  if (d_consts) cudaFree(d_consts);
  d_consts = nullptr;
  d_terms = nullptr; d_vpcs = nullptr; d_lookups = nullptr; d_code = nullptr;
}

}  // namespace zkb

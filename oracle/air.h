// ORACLE (test infrastructure, not product code).  Chips as data, the LogUp permutation trace
// and per-row quotient evaluation of the reference's multi-table STARK.
//
// The reference's AIR constraints are monomorphised Rust (`chip.eval(&mut folder)`,
// crates/stark/src/quotient.rs:157).  Both this oracle and the CUDA path consume them as DATA:
// a base-field expression DAG per chip, the shape a `SymbolicAirBuilder` walk exports
// (crates/stark/src/machine.rs:377-389) plus the chip's sends/receives as `VirtualPairCol`
// linear combinations (crates/stark/src/lookup/lookup.rs:10-19).  The wire format ("ZKMD") is
// documented in include/zkb200.h.
//
// Follows:
//   permutation trace ......... crates/stark/src/permutation.rs:16-196
//   permutation constraints ... crates/stark/src/permutation.rs:205-389
//   quotient values ........... crates/stark/src/quotient.rs:19-171
//   constraint folding ........ crates/stark/src/folder.rs:79-102 (prover), :245-249 (verifier)
//   log_quotient_degree ....... crates/stark/src/chip.rs:81-87
#pragma once
#include "pcs.h"
#include <map>

namespace zko {

enum NodeOp : u32 { N_CONST = 0, N_MAIN = 1, N_PREP = 2, N_PUB = 3, N_IS_FIRST = 4, N_IS_LAST = 5,
                    N_IS_TRANS = 6, N_ADD = 7, N_SUB = 8, N_MUL = 9, N_NEG = 10 };
struct Node { u32 op, a, b; };

struct VPC {  // VirtualPairCol: constant + sum w_i * col_i
  F constant;
  struct Term { bool is_main; u32 col; F w; };
  std::vector<Term> terms;
  template <class T, class V> T apply(const V* prep, const V* main) const {
    T r = T(constant);
    for (auto& t : terms) r = r + T(t.is_main ? main[t.col] : prep[t.col]) * t.w;
    return r;
  }
};
struct Lookup { u32 kind, scope; std::vector<VPC> values; VPC mult; };

struct Chip {
  std::string name;
  u32 prep_width = 0, main_width = 0, log_quotient_degree = 1;
  bool local_only = false, global_scope = false;
  std::vector<Lookup> sends, receives;
  std::vector<Node> nodes;
  std::vector<u32> constraints;

  // local lookups in evaluation order: sends then receives (permutation.rs:41-45)
  std::vector<std::pair<const Lookup*, bool>> local_lookups() const {
    std::vector<std::pair<const Lookup*, bool>> r;
    for (auto& l : sends) if (l.scope == 0) r.push_back({&l, true});
    for (auto& l : receives) if (l.scope == 0) r.push_back({&l, false});
    return r;
  }
  size_t batch_size() const { return (size_t)1 << log_quotient_degree; }
  size_t perm_width_ef() const {  // permutation.rs:18-23
    size_t n = local_lookups().size(), b = batch_size();
    return n == 0 ? 0 : (n + b - 1) / b + 1;
  }
  size_t num_constraints() const {  // machine.rs:377-393 + permutation.rs:355-389
    size_t c = constraints.size();
    size_t w = perm_width_ef();
    if (w) c += (w - 1) + 3;
    if (global_scope) c += 14;
    return c;
  }
};

struct Machine {
  std::vector<Chip> chips;
  u32 num_pv_elts = 0;
  FriConfig cfg;
  const Chip* find(const std::string& n) const {
    for (auto& c : chips) if (c.name == n) return &c;
    return nullptr;
  }
};

// ---- ZKMD parser --------------------------------------------------------------------------
struct WordReader {
  const u32* p; size_t n, pos = 0;
  WordReader(const u32* p_, size_t n_) : p(p_), n(n_) {}
  u32 next() { if (pos >= n) throw std::runtime_error("descriptor truncated"); return p[pos++]; }
  std::string str() {
    u32 len = next();
    std::string s;
    for (u32 i = 0; i < (len + 3) / 4; i++) { u32 w = next(); for (int b = 0; b < 4; b++) if (s.size() < len) s.push_back((char)((w >> (8 * b)) & 255)); }
    return s;
  }
};
static inline VPC parse_vpc(WordReader& r) {
  VPC v;
  v.constant = F(r.next());
  u32 nt = r.next();
  for (u32 i = 0; i < nt; i++) { VPC::Term t; t.is_main = r.next() != 0; t.col = r.next(); t.w = F(r.next()); v.terms.push_back(t); }
  return v;
}
static inline Machine parse_machine(const u32* words, size_t n) {
  WordReader r(words, n);
  if (r.next() != 0x444d4b5au) throw std::runtime_error("bad machine magic");
  if (r.next() != 1) throw std::runtime_error("bad machine version");
  Machine m;
  u32 nchips = r.next();
  m.num_pv_elts = r.next();
  m.cfg.log_blowup = r.next();
  m.cfg.num_queries = r.next();
  m.cfg.pow_bits = r.next();
  for (u32 ci = 0; ci < nchips; ci++) {
    Chip c;
    c.name = r.str();
    c.prep_width = r.next(); c.main_width = r.next(); c.log_quotient_degree = r.next();
    c.local_only = r.next() != 0; c.global_scope = r.next() != 0;
    u32 ns = r.next(), nr = r.next(), nn = r.next(), nc = r.next();
    for (u32 i = 0; i < ns + nr; i++) {
      Lookup l;
      l.kind = r.next(); l.scope = r.next();
      u32 nv = r.next();
      l.mult = parse_vpc(r);
      for (u32 j = 0; j < nv; j++) l.values.push_back(parse_vpc(r));
      (i < ns ? c.sends : c.receives).push_back(std::move(l));
    }
    for (u32 i = 0; i < nn; i++) { Node nd; nd.op = r.next(); nd.a = r.next(); nd.b = r.next(); c.nodes.push_back(nd); }
    for (u32 i = 0; i < nc; i++) c.constraints.push_back(r.next());
    if (c.log_quotient_degree > m.cfg.log_blowup) throw std::runtime_error("log_quotient_degree > log_blowup unsupported");
    m.chips.push_back(std::move(c));
  }
  return m;
}

// ---- LogUp permutation trace (permutation.rs:102-196) ---------------------------------------
// Returns the EF matrix flattened to base (n x 4E) and the local cumulative sum.
static inline Matrix generate_permutation_trace(const Chip& chip, const Matrix* prep, const Matrix& main,
                                                const E& alpha, const E& beta, E& local_sum) {
  const size_t n = main.height, Ew = chip.perm_width_ef(), B = chip.batch_size();
  Matrix out(n, 4 * Ew);
  local_sum = E::zero();
  if (Ew == 0) return out;
  auto lk = chip.local_lookups();
  size_t maxv = 0;
  for (auto& l : lk) maxv = std::max(maxv, l.first->values.size());
  std::vector<E> bpow(maxv + 1);
  { E b = E::one(); for (auto& x : bpow) { x = b; b *= beta; } }
  std::vector<E> rowsum(n);
#pragma omp parallel for schedule(static) if (n > 256)
  for (size_t r = 0; r < n; r++) {
    const F* pr = prep ? prep->row(r) : nullptr;
    const F* mr = main.row(r);
    E total;
    for (size_t b = 0; b < Ew - 1; b++) {
      E v;
      for (size_t k = b * B; k < std::min(lk.size(), (b + 1) * B); k++) {
        const Lookup& l = *lk[k].first;
        E den = alpha + bpow[0] * F(l.kind);
        for (size_t j = 0; j < l.values.size(); j++) den += bpow[j + 1] * l.values[j].apply<F, F>(pr, mr);
        F mult = l.mult.apply<F, F>(pr, mr);
        if (!lk[k].second) mult = -mult;
        v += einv(den) * mult;
      }
      for (int c = 0; c < 4; c++) out.row(r)[4 * b + c] = v.c[c];
      total += v;
    }
    rowsum[r] = total;
  }
  E run;
  for (size_t r = 0; r < n; r++) {
    run += rowsum[r];
    for (int c = 0; c < 4; c++) out.row(r)[4 * (Ew - 1) + c] = run.c[c];
  }
  local_sum = run;
  return out;
}

// ---- constraint evaluation ------------------------------------------------------------------
// One evaluator, generic over the variable type: V = F for the prover's per-row evaluation
// (base trace values), V = E for the verifier's evaluation at zeta.  Perm columns are EF.
template <class V>
struct EvalInputs {
  const V *prep_local, *prep_next, *main_local, *main_next;
  const E *perm_local, *perm_next;   // Ew entries each
  V is_first, is_last, is_trans;
  const F* pub;
  E perm_alpha, perm_beta, local_sum;
  const F* global_sum;               // 14 words
};

// calls emit(c) for every constraint in reference order; c is E.
template <class V, class Emit>
static inline void eval_constraints(const Chip& chip, const EvalInputs<V>& in, Emit emit) {
  std::vector<V> val(chip.nodes.size());
  for (size_t i = 0; i < chip.nodes.size(); i++) {
    const Node& nd = chip.nodes[i];
    switch (nd.op) {
      case N_CONST: val[i] = V(F(nd.a)); break;
      case N_MAIN: val[i] = nd.b ? in.main_next[nd.a] : in.main_local[nd.a]; break;
      case N_PREP: val[i] = nd.b ? in.prep_next[nd.a] : in.prep_local[nd.a]; break;
      case N_PUB: val[i] = V(in.pub[nd.a]); break;
      case N_IS_FIRST: val[i] = in.is_first; break;
      case N_IS_LAST: val[i] = in.is_last; break;
      case N_IS_TRANS: val[i] = in.is_trans; break;
      case N_ADD: val[i] = val[nd.a] + val[nd.b]; break;
      case N_SUB: val[i] = val[nd.a] - val[nd.b]; break;
      case N_MUL: val[i] = val[nd.a] * val[nd.b]; break;
      case N_NEG: val[i] = -val[nd.a]; break;
      default: throw std::runtime_error("bad node op");
    }
  }
  auto toE = [](const V& v) { return E(v); };
  for (u32 c : chip.constraints) emit(toE(val[c]));

  // permutation constraints (permutation.rs:205-347)
  const size_t Ew = chip.perm_width_ef(), B = chip.batch_size();
  if (Ew) {
    auto lk = chip.local_lookups();
    for (size_t b = 0; b < Ew - 1; b++) {
      std::vector<E> rlcs;
      std::vector<V> mults;
      for (size_t k = b * B; k < std::min(lk.size(), (b + 1) * B); k++) {
        const Lookup& l = *lk[k].first;
        E rlc = in.perm_alpha + F(l.kind);
        E bp = in.perm_beta;
        for (size_t j = 0; j < l.values.size(); j++) {
          V elem = l.values[j].template apply<V, V>(in.prep_local, in.main_local);
          rlc += bp * E(elem);
          bp *= in.perm_beta;
        }
        rlcs.push_back(rlc);
        V m = l.mult.template apply<V, V>(in.prep_local, in.main_local);
        mults.push_back(lk[k].second ? m : -m);
      }
      E product = E::one(), numerator = E::zero();
      for (size_t i = 0; i < rlcs.size(); i++) {
        product *= rlcs[i];
        E abc = E::one();
        for (size_t j = 0; j < rlcs.size(); j++) if (j != i) abc *= rlcs[j];
        numerator += E(mults[i]) * abc;
      }
      emit(product * in.perm_local[b] - numerator);
    }
    E sum_local, sum_next;
    for (size_t b = 0; b < Ew - 1; b++) { sum_local += in.perm_local[b]; sum_next += in.perm_next[b]; }
    E phi_local = in.perm_local[Ew - 1], phi_next = in.perm_next[Ew - 1];
    emit(E(in.is_first) * (phi_local - sum_local));
    emit(E(in.is_trans) * (phi_next - phi_local - sum_next));
    emit(E(in.is_last) * (phi_local - in.local_sum));
  }
  if (chip.global_scope) {
    const size_t M = chip.main_width;
    for (int i = 0; i < 7; i++) {
      emit(E(in.is_last) * E(in.main_local[M - 14 + i] - V(in.global_sum[i])));
      emit(E(in.is_last) * E(in.main_local[M - 7 + i] - V(in.global_sum[7 + i])));
    }
  }
}

// quotient_values (quotient.rs:19-171) on the quotient domain GENERATOR * K_{n << lqd}, natural
// order.  The three traces are given as committed bit-reversed LDEs (first Q rows = that domain).
static inline std::vector<E> quotient_values(const Chip& chip, unsigned log_n, const Matrix* prep_lde,
                                             const Matrix& main_lde, const Matrix& perm_lde,
                                             const E& perm_alpha, const E& perm_beta, const E& local_sum,
                                             const F* global_sum, const E& alpha, const F* pub) {
  const unsigned lq = log_n + chip.log_quotient_degree;
  const size_t Q = (size_t)1 << lq, n = (size_t)1 << log_n;
  const size_t C = chip.num_constraints(), Ew = chip.perm_width_ef();
  std::vector<E> apow_rev(C);   // [alpha^{C-1}, ..., alpha, 1]   prover.rs:453-456
  { E a = E::one(); for (size_t k = 0; k < C; k++) { apow_rev[C - 1 - k] = a; a *= alpha; } }
  const F gq = two_adic_generator(lq), ginv = finv(two_adic_generator(log_n));
  std::vector<F> xs(Q);
  { F c = F(GENERATOR); for (size_t i = 0; i < Q; i++) { xs[i] = c; c *= gq; } }
  std::vector<E> out(Q);
  const size_t next_step = (size_t)1 << chip.log_quotient_degree;
  std::vector<F> zero_prep(1);
#pragma omp parallel for schedule(static) if (Q > 256)
  for (size_t i = 0; i < Q; i++) {
    const size_t pl = bitrev(i, lq), pn = bitrev((i + next_step) % Q, lq);
    F x = xs[i];
    F zh = fpow(x, n) - F::one();
    EvalInputs<F> in;
    in.prep_local = prep_lde ? prep_lde->row(pl) : zero_prep.data();
    in.prep_next = prep_lde ? prep_lde->row(pn) : zero_prep.data();
    in.main_local = main_lde.row(pl);
    in.main_next = main_lde.row(pn);
    std::vector<E> pl_e(Ew), pn_e(Ew);
    for (size_t b = 0; b < Ew; b++) {
      const F* a = perm_lde.row(pl) + 4 * b; const F* c = perm_lde.row(pn) + 4 * b;
      pl_e[b] = E(a[0], a[1], a[2], a[3]); pn_e[b] = E(c[0], c[1], c[2], c[3]);
    }
    in.perm_local = pl_e.data(); in.perm_next = pn_e.data();
    in.is_first = zh * finv(x - F::one());
    in.is_last = zh * finv(x - ginv);
    in.is_trans = x - ginv;
    in.pub = pub;
    in.perm_alpha = perm_alpha; in.perm_beta = perm_beta; in.local_sum = local_sum;
    in.global_sum = global_sum;
    E acc;
    size_t k = 0;
    eval_constraints<F>(chip, in, [&](const E& c) { acc += apow_rev[k++] * c; });
    assert(k == C);
    out[i] = acc * finv(zh);
  }
  return out;
}

}  // namespace zko

// ORACLE (test infrastructure, not product code).  CPU restatement of the reference's per-shard
// protocol: StarkMachine::setup, CpuProver::{commit, open} and Verifier::verify_shard.
//
// Follows:
//   setup ............. crates/stark/src/machine.rs:352-459
//   commit ............ crates/stark/src/prover.rs:258-292
//   open .............. crates/stark/src/prover.rs:298-653   (transcript order)
//   pk.observe_into ... crates/stark/src/prover.rs:714-721
//   verify_shard ...... crates/stark/src/verifier.rs:30-246, :316-435
//   proof types ....... crates/stark/src/types.rs:16-83
// The flat proof encoding ("ZKPF") is documented in include/zkb200.h.
#pragma once
#include "air.h"

namespace zko {

// SepticDigest::zero(): CURVE_CUMULATIVE_SUM_START_{X,Y}, crates/stark/src/septic_digest.rs:9-14 (canonical)
static const u32 SEPTIC_DIGEST_ZERO[14] = {637514027u, 1595065213u, 1998064738u, 72333738u, 1211544370u, 822986770u, 1518535784u,
                                           1604177449u, 90440090u, 259343427u, 140470264u, 1162099742u, 941559812u, 1064053343u};

struct ProvingKey {
  Digest commit;
  F pc_start;
  F init_global_sum[14];
  std::vector<std::string> names;        // pk order: (height desc, name asc)
  std::vector<Matrix> traces;
  std::vector<bool> local_only;
  std::unique_ptr<CommitData> data;      // null when the machine has no preprocessed trace
  int index_of(const std::string& n) const {
    for (size_t i = 0; i < names.size(); i++) if (names[i] == n) return (int)i;
    return -1;
  }
  void observe_into(Challenger& ch) const {
    ch.observe_digest(commit);
    ch.observe(pc_start);
    ch.observe_slice(init_global_sum, 14);
    ch.observe(F::zero());
  }
};

template <class T>
static inline void sort_named(std::vector<std::pair<std::string, T>>& v) {
  std::stable_sort(v.begin(), v.end(), [](const std::pair<std::string, T>& a, const std::pair<std::string, T>& b) {
    if (a.second.height != b.second.height) return a.second.height > b.second.height;
    return a.first < b.first;
  });
}

static inline std::unique_ptr<ProvingKey> setup(const Machine& m, std::vector<std::pair<std::string, Matrix>> prep,
                                                F pc_start, const F* init_gsum) {
  auto pk = std::make_unique<ProvingKey>();
  sort_named(prep);
  std::vector<std::pair<Domain, Matrix>> in;
  for (auto& nt : prep) {
    const Chip* c = m.find(nt.first);
    if (!c) throw std::runtime_error("setup: unknown chip " + nt.first);
    if (c->prep_width != nt.second.width) throw std::runtime_error("setup: preprocessed width mismatch for " + nt.first);
    pk->names.push_back(nt.first);
    pk->local_only.push_back(c->local_only);
    in.push_back({Domain{log2_strict(nt.second.height), F::one()}, nt.second});
    pk->traces.push_back(nt.second);
  }
  if (!in.empty()) { pk->data = pcs_commit(in, m.cfg); pk->commit = pk->data->commit(); }
  else for (auto& x : pk->commit) x = F::zero();  // zero commitment, kb31_poseidon2.rs:341-345
  pk->pc_start = pc_start;
  for (int i = 0; i < 14; i++) pk->init_global_sum[i] = init_gsum[i];
  return pk;
}

struct ShardData {
  std::vector<std::string> names;      // shard order: (height desc, name asc)
  std::vector<Matrix> traces;
  std::unique_ptr<CommitData> main_data;
  std::vector<F> public_values;
};

static inline std::unique_ptr<ShardData> commit(const Machine& m, std::vector<std::pair<std::string, Matrix>> traces,
                                                const std::vector<F>& pv) {
  auto sd = std::make_unique<ShardData>();
  sort_named(traces);
  std::vector<std::pair<Domain, Matrix>> in;
  for (auto& nt : traces) {
    const Chip* c = m.find(nt.first);
    if (!c) throw std::runtime_error("commit: unknown chip " + nt.first);
    if (c->main_width != nt.second.width) throw std::runtime_error("commit: main width mismatch for " + nt.first);
    sd->names.push_back(nt.first);
    in.push_back({Domain{log2_strict(nt.second.height), F::one()}, nt.second});
    sd->traces.push_back(std::move(nt.second));
  }
  sd->main_data = pcs_commit(in, m.cfg);
  sd->public_values = pv;
  return sd;
}

struct AirOpened { std::vector<E> local, next; };
struct ChipOpened {
  std::string name;
  AirOpened preprocessed, main, permutation;   // permutation: 4E "EF of base columns" entries
  std::vector<std::array<E, 4>> quotient;       // per chunk, 4 EF values
  F global_sum[14];
  E local_sum;
  u32 log_degree;
};
struct ShardProof {
  Digest main_commit, perm_commit, quot_commit;
  std::vector<ChipOpened> chips;
  std::vector<F> public_values;
  FriProof fri;
};

static inline std::unique_ptr<ShardProof> open(const Machine& m, const ProvingKey& pk, ShardData& sd, Challenger& ch) {
  auto proof = std::make_unique<ShardProof>();
  const size_t nc = sd.names.size();
  std::vector<const Chip*> chips;
  std::vector<unsigned> logn;
  for (size_t i = 0; i < nc; i++) { chips.push_back(m.find(sd.names[i])); logn.push_back(log2_strict(sd.traces[i].height)); }

  ch.observe_slice(sd.public_values.data(), m.num_pv_elts);
  ch.observe_digest(sd.main_data->commit());
  E perm_alpha = ch.sample_ext(), perm_beta = ch.sample_ext();

  // permutation traces
  std::vector<E> local_sums(nc);
  std::vector<std::array<F, 14>> global_sums(nc);
  std::vector<std::pair<Domain, Matrix>> perm_in;
  for (size_t i = 0; i < nc; i++) {
    int pi = pk.index_of(sd.names[i]);
    const Matrix* prep = pi >= 0 ? &pk.traces[pi] : nullptr;
    if (prep && prep->height != sd.traces[i].height) throw std::runtime_error("preprocessed and main have different heights");
    Matrix pt = generate_permutation_trace(*chips[i], prep, sd.traces[i], perm_alpha, perm_beta, local_sums[i]);
    // Local-scope chips carry SepticDigest::zero(), the curve START point, NOT 14 zeros
    // (crates/stark/src/septic_digest.rs:9-42, prover.rs:352)
    for (int k = 0; k < 14; k++) global_sums[i][k] = F(SEPTIC_DIGEST_ZERO[k]);
    if (chips[i]->global_scope) {
      const Matrix& t = sd.traces[i];
      const F* last = t.row(t.height - 1) + t.width - 14;
      for (int k = 0; k < 14; k++) global_sums[i][k] = last[k];
    }
    perm_in.push_back({Domain{logn[i], F::one()}, std::move(pt)});
  }
  auto perm_data = pcs_commit(perm_in, m.cfg);
  ch.observe_digest(perm_data->commit());
  for (size_t i = 0; i < nc; i++) {
    ch.observe_ext(local_sums[i]);
    ch.observe_slice(global_sums[i].data(), 14);
  }

  // quotient
  const E alpha = ch.sample_ext();
  std::vector<std::pair<Domain, Matrix>> quot_in;
  for (size_t i = 0; i < nc; i++) {
    int pi = pk.index_of(sd.names[i]);
    const Matrix* prep_lde = pi >= 0 ? &pk.data->tree.mats[pi] : nullptr;
    std::vector<E> q = quotient_values(*chips[i], logn[i], prep_lde, sd.main_data->tree.mats[i], perm_data->tree.mats[i],
                                       perm_alpha, perm_beta, local_sums[i], global_sums[i].data(), alpha,
                                       sd.public_values.data());
    const unsigned lqd = chips[i]->log_quotient_degree;
    const size_t nchunks = (size_t)1 << lqd, n = (size_t)1 << logn[i];
    const F gq = two_adic_generator(logn[i] + lqd);
    for (size_t j = 0; j < nchunks; j++) {
      Matrix chunk(n, 4);
      for (size_t k = 0; k < n; k++) for (int c = 0; c < 4; c++) chunk.row(k)[c] = q[k * nchunks + j].c[c];
      quot_in.push_back({Domain{logn[i], F(GENERATOR) * fpow(gq, j)}, std::move(chunk)});
    }
  }
  auto quot_data = pcs_commit(quot_in, m.cfg);
  ch.observe_digest(quot_data->commit());

  const E zeta = ch.sample_ext();
  std::vector<OpenRound> rounds;
  OpenRound r_prep{pk.data.get(), {}}, r_main{sd.main_data.get(), {}}, r_perm{perm_data.get(), {}}, r_quot{quot_data.get(), {}};
  for (size_t i = 0; i < pk.traces.size(); i++) {
    E nxt = zeta * two_adic_generator(log2_strict(pk.traces[i].height));
    r_prep.points.push_back(pk.local_only[i] ? std::vector<E>{zeta} : std::vector<E>{zeta, nxt});
  }
  for (size_t i = 0; i < nc; i++) {
    E nxt = zeta * two_adic_generator(logn[i]);
    r_main.points.push_back(chips[i]->local_only ? std::vector<E>{zeta} : std::vector<E>{zeta, nxt});
    r_perm.points.push_back({zeta, nxt});
  }
  for (size_t i = 0; i < quot_in.size(); i++) r_quot.points.push_back({zeta});
  if (pk.data) rounds.push_back(r_prep);
  rounds.push_back(r_main); rounds.push_back(r_perm); rounds.push_back(r_quot);
  std::vector<RoundOpenings> opened;
  pcs_open(rounds, ch, m.cfg, opened, proof->fri);
  const size_t ro = pk.data ? 1 : 0;   // index of the main round

  proof->main_commit = sd.main_data->commit();
  proof->perm_commit = perm_data->commit();
  proof->quot_commit = quot_data->commit();
  proof->public_values = sd.public_values;
  size_t qoff = 0;
  for (size_t i = 0; i < nc; i++) {
    ChipOpened co;
    co.name = sd.names[i];
    co.log_degree = logn[i];
    auto fill = [](AirOpened& a, const std::vector<std::vector<E>>& pts) {
      a.local = pts[0];
      a.next = pts.size() > 1 ? pts[1] : std::vector<E>(pts[0].size(), E::zero());
    };
    int pi = pk.index_of(sd.names[i]);
    if (pi >= 0) fill(co.preprocessed, opened[0][pi]);
    fill(co.main, opened[ro][i]);
    fill(co.permutation, opened[ro + 1][i]);
    const size_t nchunks = (size_t)1 << chips[i]->log_quotient_degree;
    for (size_t j = 0; j < nchunks; j++) {
      const std::vector<E>& v = opened[ro + 2][qoff + j][0];
      co.quotient.push_back({v[0], v[1], v[2], v[3]});
    }
    qoff += nchunks;
    for (int k = 0; k < 14; k++) co.global_sum[k] = global_sums[i][k];
    co.local_sum = local_sums[i];
    proof->chips.push_back(std::move(co));
  }
  return proof;
}

// ---- flat proof encoding -----------------------------------------------------------------------
struct WordWriter {
  std::vector<u32> w;
  void put(u32 x) { w.push_back(x); }
  void put(F x) { w.push_back(x.v); }
  void put(const E& e) { for (int i = 0; i < 4; i++) w.push_back(e.c[i].v); }
  void put(const Digest& d) { for (auto& x : d) w.push_back(x.v); }
  void str(const std::string& s) {
    put((u32)s.size());
    for (size_t i = 0; i < s.size(); i += 4) { u32 x = 0; for (size_t b = 0; b < 4 && i + b < s.size(); b++) x |= (u32)(unsigned char)s[i + b] << (8 * b); put(x); }
  }
};
static inline std::vector<u32> serialize_proof(const ShardProof& p) {
  WordWriter o;
  o.put(0x46504b5au); o.put((u32)1);
  o.put(p.main_commit); o.put(p.perm_commit); o.put(p.quot_commit);
  o.put((u32)p.chips.size());
  for (auto& c : p.chips) {
    o.str(c.name);
    o.put(c.log_degree);
    o.put((u32)c.preprocessed.local.size()); o.put((u32)c.main.local.size());
    o.put((u32)c.permutation.local.size()); o.put((u32)c.quotient.size());
    for (auto* a : {&c.preprocessed, &c.main, &c.permutation}) { for (auto& e : a->local) o.put(e); for (auto& e : a->next) o.put(e); }
    for (auto& q : c.quotient) for (auto& e : q) o.put(e);
    for (int k = 0; k < 14; k++) o.put(c.global_sum[k]);
    o.put(c.local_sum);
  }
  o.put((u32)p.public_values.size());
  for (auto& x : p.public_values) o.put(x);
  const FriProof& f = p.fri;
  o.put((u32)f.commit_phase_commits.size());
  for (auto& d : f.commit_phase_commits) o.put(d);
  o.put(f.final_poly); o.put(f.pow_witness);
  o.put((u32)f.query_proofs.size());
  for (auto& q : f.query_proofs) {
    o.put((u32)q.input_proof.size());
    for (auto& b : q.input_proof) {
      o.put((u32)b.opened_values.size());
      for (auto& r : b.opened_values) { o.put((u32)r.size()); for (auto& x : r) o.put(x); }
      o.put((u32)b.opening_proof.size());
      for (auto& d : b.opening_proof) o.put(d);
    }
    o.put((u32)q.commit_phase_openings.size());
    for (auto& s : q.commit_phase_openings) {
      o.put(s.sibling_value);
      o.put((u32)s.opening_proof.size());
      for (auto& d : s.opening_proof) o.put(d);
    }
  }
  return o.w;
}
static inline ShardProof parse_proof(const u32* words, size_t n) {
  WordReader r(words, n);
  auto f = [&]() { return F(r.next()); };
  auto e = [&]() { E x; for (int i = 0; i < 4; i++) x.c[i] = F(r.next()); return x; };
  auto dg = [&]() { Digest d; for (auto& x : d) x = F(r.next()); return d; };
  if (r.next() != 0x46504b5au || r.next() != 1) throw std::runtime_error("bad proof header");
  ShardProof p;
  p.main_commit = dg(); p.perm_commit = dg(); p.quot_commit = dg();
  u32 nc = r.next();
  for (u32 i = 0; i < nc; i++) {
    ChipOpened c;
    c.name = r.str();
    c.log_degree = r.next();
    u32 pw = r.next(), mw = r.next(), ew = r.next(), nq = r.next();
    u32 ws[3] = {pw, mw, ew};
    AirOpened* as[3] = {&c.preprocessed, &c.main, &c.permutation};
    for (int k = 0; k < 3; k++) { for (u32 j = 0; j < ws[k]; j++) as[k]->local.push_back(e()); for (u32 j = 0; j < ws[k]; j++) as[k]->next.push_back(e()); }
    for (u32 j = 0; j < nq; j++) { std::array<E, 4> q; for (auto& x : q) x = e(); c.quotient.push_back(q); }
    for (int k = 0; k < 14; k++) c.global_sum[k] = f();
    c.local_sum = e();
    p.chips.push_back(std::move(c));
  }
  u32 npv = r.next();
  for (u32 i = 0; i < npv; i++) p.public_values.push_back(f());
  u32 ncm = r.next();
  for (u32 i = 0; i < ncm; i++) p.fri.commit_phase_commits.push_back(dg());
  p.fri.final_poly = e(); p.fri.pow_witness = f();
  u32 nq = r.next();
  for (u32 i = 0; i < nq; i++) {
    QueryProof q;
    u32 nr = r.next();
    for (u32 j = 0; j < nr; j++) {
      BatchOpening b;
      u32 nm = r.next();
      for (u32 k = 0; k < nm; k++) { u32 w = r.next(); std::vector<F> row; for (u32 t = 0; t < w; t++) row.push_back(f()); b.opened_values.push_back(std::move(row)); }
      u32 pl = r.next();
      for (u32 k = 0; k < pl; k++) b.opening_proof.push_back(dg());
      q.input_proof.push_back(std::move(b));
    }
    u32 nl = r.next();
    for (u32 j = 0; j < nl; j++) {
      CommitPhaseStep s;
      s.sibling_value = e();
      u32 pl = r.next();
      for (u32 k = 0; k < pl; k++) s.opening_proof.push_back(dg());
      q.commit_phase_openings.push_back(std::move(s));
    }
    p.fri.query_proofs.push_back(std::move(q));
  }
  return p;
}

// ---- verifier (verifier.rs:30-246; public values observed by the caller, machine.rs:639-641) ----
struct VerifyingKey {
  Digest commit; F pc_start; F init_global_sum[14];
  std::vector<std::string> names; std::vector<unsigned> log_n; std::vector<u32> widths;   // chip_information
  void observe_into(Challenger& ch) const {
    ch.observe_digest(commit); ch.observe(pc_start); ch.observe_slice(init_global_sum, 14); ch.observe(F::zero());
  }
};
static inline VerifyingKey vk_from_pk(const ProvingKey& pk) {
  VerifyingKey vk;
  vk.commit = pk.commit; vk.pc_start = pk.pc_start;
  for (int i = 0; i < 14; i++) vk.init_global_sum[i] = pk.init_global_sum[i];
  vk.names = pk.names;
  for (auto& t : pk.traces) { vk.log_n.push_back(log2_strict(t.height)); vk.widths.push_back((u32)t.width); }
  return vk;
}

static inline std::string verify_shard(const Machine& m, const VerifyingKey& vk, Challenger& ch, const ShardProof& proof) {
  const size_t nc = proof.chips.size();
  std::vector<const Chip*> chips;
  for (auto& c : proof.chips) { const Chip* ch_ = m.find(c.name); if (!ch_) return "unknown chip " + c.name; chips.push_back(ch_); }
  if (proof.public_values.size() < m.num_pv_elts) return "public values too short";
  ch.observe_slice(proof.public_values.data(), m.num_pv_elts);
  ch.observe_digest(proof.main_commit);
  E perm_alpha = ch.sample_ext(), perm_beta = ch.sample_ext();
  ch.observe_digest(proof.perm_commit);
  for (size_t i = 0; i < nc; i++) {
    const ChipOpened& o = proof.chips[i];
    ch.observe_ext(o.local_sum);
    ch.observe_slice(o.global_sum, 14);
    bool gz = true;
    for (int k = 0; k < 14; k++) gz = gz && o.global_sum[k] == F(SEPTIC_DIGEST_ZERO[k]);   // SepticDigest::is_zero, verifier.rs:102
    if (!chips[i]->global_scope && !gz) return "global cumulative sum is non-zero, but chip is Local";
    if (chips[i]->perm_width_ef() == 0 && !o.local_sum.is_zero()) return "local cumulative sum is non-zero, but no local lookups";
  }
  const E alpha = ch.sample_ext();
  ch.observe_digest(proof.quot_commit);
  const E zeta = ch.sample_ext();

  auto index_of = [&](const std::string& n) { for (size_t i = 0; i < nc; i++) if (proof.chips[i].name == n) return (int)i; return -1; };
  std::vector<VerifyRound> rounds;
  if (!vk.names.empty()) {
    VerifyRound r; r.commit = vk.commit;
    for (size_t k = 0; k < vk.names.size(); k++) {
      int i = index_of(vk.names[k]);
      if (i < 0) return "preprocessed chip " + vk.names[k] + " missing from shard";
      VerifyMat vm; vm.domain = Domain{vk.log_n[k], F::one()};
      vm.points_and_values.push_back({zeta, proof.chips[i].preprocessed.local});
      if (!chips[i]->local_only) vm.points_and_values.push_back({zeta * two_adic_generator(vk.log_n[k]), proof.chips[i].preprocessed.next});
      r.mats.push_back(std::move(vm));
    }
    rounds.push_back(std::move(r));
  }
  VerifyRound rm, rp, rq;
  rm.commit = proof.main_commit; rp.commit = proof.perm_commit; rq.commit = proof.quot_commit;
  for (size_t i = 0; i < nc; i++) {
    const ChipOpened& o = proof.chips[i];
    E nxt = zeta * two_adic_generator(o.log_degree);
    VerifyMat a; a.domain = Domain{o.log_degree, F::one()};
    a.points_and_values.push_back({zeta, o.main.local});
    if (!chips[i]->local_only) a.points_and_values.push_back({nxt, o.main.next});
    rm.mats.push_back(std::move(a));
    VerifyMat b; b.domain = Domain{o.log_degree, F::one()};
    b.points_and_values.push_back({zeta, o.permutation.local});
    b.points_and_values.push_back({nxt, o.permutation.next});
    rp.mats.push_back(std::move(b));
    const unsigned lqd = chips[i]->log_quotient_degree;
    const F gq = two_adic_generator(o.log_degree + lqd);
    if (o.quotient.size() != ((size_t)1 << lqd)) return "quotient chunk count";
    for (size_t j = 0; j < o.quotient.size(); j++) {
      VerifyMat c; c.domain = Domain{o.log_degree, F(GENERATOR) * fpow(gq, j)};
      c.points_and_values.push_back({zeta, std::vector<E>(o.quotient[j].begin(), o.quotient[j].end())});
      rq.mats.push_back(std::move(c));
    }
  }
  rounds.push_back(std::move(rm)); rounds.push_back(std::move(rp)); rounds.push_back(std::move(rq));
  std::string err = pcs_verify(rounds, proof.fri, ch, m.cfg);
  if (!err.empty()) return "invalid opening argument: " + err;

  E total_local;
  for (size_t i = 0; i < nc; i++) {
    const Chip& chip = *chips[i];
    const ChipOpened& o = proof.chips[i];
    // opening shape (verifier.rs:248-312)
    if (o.preprocessed.local.size() != chip.prep_width || o.preprocessed.next.size() != chip.prep_width) return "preprocessed width mismatch on " + chip.name;
    if (o.main.local.size() != chip.main_width || o.main.next.size() != chip.main_width) return "main width mismatch on " + chip.name;
    if (o.permutation.local.size() != 4 * chip.perm_width_ef() || o.permutation.next.size() != 4 * chip.perm_width_ef()) return "permutation width mismatch on " + chip.name;
    // recompute quotient(zeta) from the chunks (verifier.rs:399-435)
    const unsigned lqd = chip.log_quotient_degree;
    const size_t nch = (size_t)1 << lqd;
    const F gq = two_adic_generator(o.log_degree + lqd);
    std::vector<F> shifts(nch);
    for (size_t j = 0; j < nch; j++) shifts[j] = F(GENERATOR) * fpow(gq, j);
    auto zp_at = [&](F shift, const E& pt) { E u = pt * finv(shift); for (unsigned k = 0; k < o.log_degree; k++) u *= u; return u - F::one(); };
    E quotient;
    for (size_t a = 0; a < nch; a++) {
      E zps = E::one();
      for (size_t b = 0; b < nch; b++) if (b != a) zps *= zp_at(shifts[b], zeta) * einv(zp_at(shifts[b], E(shifts[a])));
      E inner;
      for (int e_i = 0; e_i < 4; e_i++) { E mono; mono.c[e_i] = F::one(); inner += mono * o.quotient[a][e_i]; }
      quotient += zps * inner;
    }
    // constraints at zeta (verifier.rs:353-397)
    SelectorsE sel = selectors_at_point(o.log_degree, zeta);
    auto unflatten = [&](const std::vector<E>& v) {
      std::vector<E> r;
      for (size_t k = 0; k + 3 < v.size(); k += 4) { E s; for (int e_i = 0; e_i < 4; e_i++) { E mono; mono.c[e_i] = F::one(); s += mono * v[k + e_i]; } r.push_back(s); }
      return r;
    };
    std::vector<E> pl = unflatten(o.permutation.local), pn = unflatten(o.permutation.next);
    std::vector<E> zero1(1);
    EvalInputs<E> in;
    in.prep_local = chip.prep_width ? o.preprocessed.local.data() : zero1.data();
    in.prep_next = chip.prep_width ? o.preprocessed.next.data() : zero1.data();
    in.main_local = o.main.local.data(); in.main_next = o.main.next.data();
    in.perm_local = pl.data(); in.perm_next = pn.data();
    in.is_first = sel.is_first_row; in.is_last = sel.is_last_row; in.is_trans = sel.is_transition;
    in.pub = proof.public_values.data();
    in.perm_alpha = perm_alpha; in.perm_beta = perm_beta; in.local_sum = o.local_sum;
    in.global_sum = o.global_sum;
    E acc;
    eval_constraints<E>(chip, in, [&](const E& c) { acc = acc * alpha + c; });
    if (acc * sel.inv_zeroifier != quotient) return "out-of-domain evaluation mismatch on chip " + chip.name;
    total_local += o.local_sum;
  }
  if (!total_local.is_zero()) return "local cumulative sum is not zero";
  return "";
}

}  // namespace zko

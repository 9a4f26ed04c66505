// ORACLE (test infrastructure, not product code).  CPU restatement of the polynomial
// commitment scheme the reference's shard prover drives: radix-2 DFT / coset LDE, the
// Poseidon2 Merkle-tree MMCS with mixed heights, and TwoAdicFriPcs {commit, open, verify}.
//
// The arithmetic itself lives in the un-vendored dependency ProjectZKM/Plonky3 @
// faa24ca4597eebeecbf71b194b71c7d1a99b3f01 (Cargo.lock:4014-4320); this file restates its
// published algorithm and is anchored on the in-tree verifier restatement:
//   TwoAdicFriPcs::verify ......... crates/recursion/circuit/src/fri.rs:71-218
//   FRI challenges / shape ........ crates/recursion/circuit/src/fri.rs:34-69
//   FRI query fold ................ crates/recursion/circuit/src/fri.rs:247-361
//   MMCS verify_batch ............. crates/recursion/circuit/src/fri.rs:363-405
//   proof shapes .................. crates/recursion/circuit/src/fri.rs:411-465
//   call sites of commit/open ..... crates/stark/src/prover.rs:277,403,497,546-556
// PARITY NOTE: the reference holds no golden commitments or proofs for this path (its tests
// use random inputs and check verify()==Ok), so LDE/Merkle/FRI outputs are pinned by
// mathematical uniqueness + acceptance by the verifier restated below, not by golden values.
#pragma once
#include "hash.h"
#include <algorithm>
#include <list>
#include <memory>
#include <stdexcept>
#include <string>

namespace zko {

struct Matrix {  // row-major, like the reference's RowMajorMatrix
  size_t height = 0, width = 0;
  std::vector<F> v;
  Matrix() {}
  Matrix(size_t h, size_t w) : height(h), width(w), v(h * w) {}
  F* row(size_t r) { return v.data() + r * width; }
  const F* row(size_t r) const { return v.data() + r * width; }
};

// In-place radix-2 DIT over the rows of a matrix (all columns at once), natural in/out.
static inline void dft_rows(Matrix& m, bool inverse) {
  const size_t n = m.height, w = m.width;
  if (n <= 1) return;
  const unsigned ln = log2_strict(n);
  for (size_t i = 0; i < n; i++) {
    size_t j = bitrev(i, ln);
    if (i < j) std::swap_ranges(m.row(i), m.row(i) + w, m.row(j));
  }
  F root = two_adic_generator(ln);
  if (inverse) root = finv(root);
  std::vector<F> tw(n / 2);
  tw[0] = F::one();
  for (size_t i = 1; i < n / 2; i++) tw[i] = tw[i - 1] * root;
  for (unsigned s = 0; s < ln; s++) {
    const size_t half = (size_t)1 << s, step = n >> (s + 1);
#pragma omp parallel for schedule(static) if (n * w > (1u << 14))
    for (size_t b = 0; b < n / 2; b++) {
      size_t blk = b >> s, k = b & (half - 1);
      size_t i0 = (blk << (s + 1)) + k, i1 = i0 + half;
      F t = tw[k * step];
      F* r0 = m.row(i0);
      F* r1 = m.row(i1);
      for (size_t c = 0; c < w; c++) {
        F u = r0[c], x = r1[c] * t;
        r0[c] = u + x;
        r1[c] = u - x;
      }
    }
  }
  if (inverse) {
    F ninv = finv(F((u32)n));
#pragma omp parallel for schedule(static) if (n * w > (1u << 14))
    for (size_t i = 0; i < n * w; i++) m.v[i] *= ninv;
  }
}

// coset_lde_batch(evals over H_n, added_bits, shift) followed by bit_reverse_rows:
// result row bitrev(i) = p(shift * w_{n<<added}^i).   [P3-upstream TwoAdicFriPcs::commit]
static inline Matrix coset_lde_bitrev(const Matrix& evals, unsigned added_bits, F shift) {
  const size_t n = evals.height, w = evals.width, N = n << added_bits;
  Matrix c = evals;
  dft_rows(c, true);
  Matrix big(N, w);
  F s = F::one();
  for (size_t i = 0; i < n; i++) {
    for (size_t j = 0; j < w; j++) big.row(i)[j] = c.row(i)[j] * s;
    s *= shift;
  }
  dft_rows(big, false);
  Matrix out(N, w);
  const unsigned lN = log2_strict(N);
#pragma omp parallel for schedule(static) if (N * w > (1u << 14))
  for (size_t i = 0; i < N; i++) {
    const F* src = big.row(i);
    F* dst = out.row(bitrev(i, lN));
    for (size_t j = 0; j < w; j++) dst[j] = src[j];
  }
  return out;
}

// ---- MMCS (MerkleTreeMmcs<.., 8> over PaddingFreeSponge / TruncatedPermutation) -----------
struct MerkleTree {
  std::vector<Matrix> mats;                   // committed matrices, caller's order
  std::vector<std::vector<Digest>> layers;    // layers[0] = leaf digests of the tallest
  Digest root;
  size_t max_height = 0;

  static Digest hash_rows(const std::vector<const Matrix*>& ms, size_t r) {
    std::vector<F> buf;
    for (auto* m : ms) buf.insert(buf.end(), m->row(r), m->row(r) + m->width);
    return sponge_hash(buf.data(), buf.size());
  }
  void build() {
    max_height = 0;
    for (auto& m : mats) max_height = std::max(max_height, m.height);
    auto at_height = [&](size_t h) {
      std::vector<const Matrix*> r;
      for (auto& m : mats) if (m.height == h) r.push_back(&m);
      return r;
    };
    layers.clear();
    auto tall = at_height(max_height);
    std::vector<Digest> cur(max_height);
#pragma omp parallel for schedule(static) if (max_height > 64)
    for (size_t r = 0; r < max_height; r++) cur[r] = hash_rows(tall, r);
    layers.push_back(cur);
    while (layers.back().size() > 1) {
      const std::vector<Digest>& prev = layers.back();
      size_t n = prev.size() / 2;
      auto inj = at_height(n);
      std::vector<Digest> next(n);
#pragma omp parallel for schedule(static) if (n > 64)
      for (size_t i = 0; i < n; i++) {
        Digest d = compress2(prev[2 * i], prev[2 * i + 1]);
        if (!inj.empty()) d = compress2(d, hash_rows(inj, i));
        next[i] = d;
      }
      layers.push_back(std::move(next));
    }
    root = layers.back()[0];
  }
  unsigned log_max_height() const { return log2_strict(max_height); }
  // open_batch: one row per matrix (index shifted down for shorter ones) + sibling path.
  void open_batch(size_t index, std::vector<std::vector<F>>& rows, std::vector<Digest>& path) const {
    unsigned lm = log_max_height();
    rows.clear();
    for (auto& m : mats) {
      size_t r = index >> (lm - log2_strict(m.height));
      rows.emplace_back(m.row(r), m.row(r) + m.width);
    }
    path.clear();
    for (unsigned l = 0; l < lm; l++) path.push_back(layers[l][(index >> l) ^ 1]);
  }
};

// crates/recursion/circuit/src/fri.rs:363-405
static inline bool verify_batch(const Digest& commit, const std::vector<size_t>& dims, size_t index,
                                unsigned n_index_bits, const std::vector<std::vector<F>>& opened,
                                const std::vector<Digest>& proof) {
  std::vector<size_t> order(dims.size());
  for (size_t i = 0; i < order.size(); i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return dims[a] > dims[b]; });
  size_t pos = 0, cur = dims[order[0]];
  auto take = [&](size_t h) {
    std::vector<F> buf;
    while (pos < order.size() && dims[order[pos]] == h) {
      buf.insert(buf.end(), opened[order[pos]].begin(), opened[order[pos]].end());
      pos++;
    }
    return buf;
  };
  std::vector<F> buf = take(cur);
  Digest root = sponge_hash(buf.data(), buf.size());
  if (proof.size() != n_index_bits) return false;
  for (unsigned l = 0; l < n_index_bits; l++) {
    bool bit = (index >> l) & 1;
    root = bit ? compress2(proof[l], root) : compress2(root, proof[l]);
    cur >>= 1;
    if (pos < order.size() && dims[order[pos]] == cur) {
      std::vector<F> b2 = take(cur);
      root = compress2(root, sponge_hash(b2.data(), b2.size()));
    }
  }
  return root == commit && pos == order.size();
}

// ---- TwoAdicFriPcs -------------------------------------------------------------------------
struct FriConfig {
  unsigned log_blowup = 1, num_queries = 84, pow_bits = 16;  // crates/stark/src/kb31_poseidon2.rs:203-213
};

// domain: coset shift * H_{2^log_n}
struct Domain { unsigned log_n; F shift; };

struct CommitData {
  MerkleTree tree;  // tree.mats = bit-reversed LDEs on GENERATOR * K
  Digest commit() const { return tree.root; }
};

static inline std::unique_ptr<CommitData> pcs_commit(const std::vector<std::pair<Domain, Matrix>>& in,
                                                     const FriConfig& cfg) {
  auto cd = std::make_unique<CommitData>();
  for (auto& dm : in) {
    assert(((size_t)1 << dm.first.log_n) == dm.second.height);
    F shift = F(GENERATOR) * finv(dm.first.shift);
    cd->tree.mats.push_back(coset_lde_bitrev(dm.second, cfg.log_blowup, shift));
  }
  cd->tree.build();
  return cd;
}

struct BatchOpening { std::vector<std::vector<F>> opened_values; std::vector<Digest> opening_proof; };
struct CommitPhaseStep { E sibling_value; std::vector<Digest> opening_proof; };
struct QueryProof { std::vector<BatchOpening> input_proof; std::vector<CommitPhaseStep> commit_phase_openings; };
struct FriProof {
  std::vector<Digest> commit_phase_commits;
  std::vector<QueryProof> query_proofs;
  E final_poly;
  F pow_witness;
};

static inline void batch_inverse(std::vector<E>& v) {
  std::vector<E> pre(v.size());
  E acc = E::one();
  for (size_t i = 0; i < v.size(); i++) { pre[i] = acc; acc *= v[i]; }
  E inv = einv(acc);
  for (size_t i = v.size(); i-- > 0;) { E t = inv * pre[i]; inv *= v[i]; v[i] = t; }
}

// p(z) for every column, from the n evaluations on GENERATOR * H_n that sit in the first n rows
// of the bit-reversed LDE.
static inline std::vector<E> interpolate_low_coset(const Matrix& lde, unsigned log_blowup, const E& z) {
  const size_t n = lde.height >> log_blowup, w = lde.width;
  const unsigned ln = log2_strict(n);
  E zp = z * finv(F(GENERATOR));
  E zn = zp;
  for (unsigned i = 0; i < ln; i++) zn *= zn;
  E scale = (zn - F::one()) * finv(F((u32)n));
  F g = two_adic_generator(ln);
  std::vector<F> h(n);
  h[0] = F::one();
  for (size_t i = 1; i < n; i++) h[i] = h[i - 1] * g;
  std::vector<E> den(n);
  for (size_t i = 0; i < n; i++) den[i] = zp - h[i];
  batch_inverse(den);
  std::vector<E> wgt(n);
  for (size_t pos = 0; pos < n; pos++) { size_t i = bitrev(pos, ln); wgt[pos] = den[i] * h[i]; }
  std::vector<E> ys(w);
#pragma omp parallel for schedule(static) if (n * w > (1u << 14))
  for (size_t cb = 0; cb < (w + 15) / 16; cb++) {
    size_t c0 = cb * 16, c1 = std::min(w, c0 + 16);
    for (size_t pos = 0; pos < n; pos++) {
      const F* r = lde.row(pos);
      for (size_t c = c0; c < c1; c++) ys[c] += wgt[pos] * r[c];
    }
  }
  for (size_t c = 0; c < w; c++) ys[c] *= scale;
  return ys;
}

typedef std::vector<std::vector<std::vector<E>>> RoundOpenings;  // [mat][point][col]
struct OpenRound { const CommitData* data; std::vector<std::vector<E>> points; };  // points[mat]

static inline E fold_pair(const E& e0, const E& e1, F x0, const E& beta) {
  // e0 + (beta - x0)(e1 - e0)/(x1 - x0), x1 = -x0      crates/recursion/circuit/src/fri.rs:308-340
  F inv = finv(-(x0 + x0));
  return e0 + (beta - x0) * ((e1 - e0) * inv);
}

static inline void pcs_open(const std::vector<OpenRound>& rounds, Challenger& ch, const FriConfig& cfg,
                            std::vector<RoundOpenings>& all_opened, FriProof& proof) {
  const E alpha = ch.sample_ext();
  unsigned log_gmax = 0;
  for (auto& r : rounds) log_gmax = std::max(log_gmax, r.data->tree.log_max_height());

  std::vector<std::vector<E>> ro(32);
  size_t num_reduced[32] = {0};
  struct DenCache { unsigned lh; E z; std::vector<E> den; };
  std::list<DenCache> den_cache;
  all_opened.clear();
  for (auto& r : rounds) {
    RoundOpenings ropen;
    for (size_t mi = 0; mi < r.data->tree.mats.size(); mi++) {
      const Matrix& lde = r.data->tree.mats[mi];
      const unsigned lh = log2_strict(lde.height);
      if (ro[lh].empty()) ro[lh].assign(lde.height, E::zero());
      // reduced_row(x) = sum_j alpha^j p_j(x)
      std::vector<E> apow(lde.width);
      { E a = E::one(); for (size_t j = 0; j < lde.width; j++) { apow[j] = a; a *= alpha; } }
      std::vector<E> rrow(lde.height);
#pragma omp parallel for schedule(static) if (lde.height * lde.width > (1u << 14))
      for (size_t x = 0; x < lde.height; x++) {
        E acc;
        const F* row = lde.row(x);
        for (size_t j = 0; j < lde.width; j++) acc += apow[j] * row[j];
        rrow[x] = acc;
      }
      std::vector<std::vector<E>> per_point;
      for (const E& z : r.points[mi]) {
        std::vector<E> ys = interpolate_low_coset(lde, cfg.log_blowup, z);
        E rys;
        for (size_t j = 0; j < lde.width; j++) rys += apow[j] * ys[j];
        E off = epow(alpha, num_reduced[lh]);
        // 1/(z - x), x = GENERATOR * w^{bitrev(pos)}; shared by every matrix of this height
        const std::vector<E>* denp = nullptr;
        for (auto& c : den_cache) if (c.lh == lh && c.z == z) denp = &c.den;
        if (!denp) {
          DenCache dc; dc.lh = lh; dc.z = z; dc.den.resize(lde.height);
          F g = two_adic_generator(lh);
          { F c = F(GENERATOR); for (size_t i = 0; i < lde.height; i++) { dc.den[bitrev(i, lh)] = z - c; c *= g; } }
          batch_inverse(dc.den);
          den_cache.push_back(std::move(dc));
          denp = &den_cache.back().den;
        }
        const std::vector<E>& den = *denp;
#pragma omp parallel for schedule(static) if (lde.height > (1u << 12))
        for (size_t x = 0; x < lde.height; x++) ro[lh][x] += off * (rys - rrow[x]) * den[x];
        num_reduced[lh] += lde.width;
        per_point.push_back(std::move(ys));
      }
      ropen.push_back(std::move(per_point));
    }
    all_opened.push_back(std::move(ropen));
  }

  // ---- FRI commit phase ------------------------------------------------------------------
  std::vector<E> folded = ro[log_gmax];
  std::vector<std::unique_ptr<MerkleTree>> layer_trees;
  proof.commit_phase_commits.clear();
  const size_t blowup = (size_t)1 << cfg.log_blowup;
  while (folded.size() > blowup) {
    const size_t m = folded.size();
    const unsigned lm = log2_strict(m);
    auto tree = std::make_unique<MerkleTree>();
    Matrix leaves(m / 2, 8);
    for (size_t i = 0; i < m / 2; i++)
      for (int k = 0; k < 2; k++)
        for (int c = 0; c < 4; c++) leaves.row(i)[4 * k + c] = folded[2 * i + k].c[c];
    tree->mats.push_back(std::move(leaves));
    tree->build();
    ch.observe_digest(tree->root);
    proof.commit_phase_commits.push_back(tree->root);
    const E beta = ch.sample_ext();
    std::vector<E> next(m / 2);
    F g = two_adic_generator(lm);
    for (size_t i = 0; i < m / 2; i++) {
      F x0 = fpow(g, bitrev(2 * i, lm));
      next[i] = fold_pair(folded[2 * i], folded[2 * i + 1], x0, beta);
    }
    const unsigned lnext = lm - 1;
    if (!ro[lnext].empty()) {
      E b2 = beta * beta;
      for (size_t i = 0; i < next.size(); i++) next[i] += b2 * ro[lnext][i];
    }
    folded = std::move(next);
    layer_trees.push_back(std::move(tree));
  }
  proof.final_poly = folded[0];
  for (auto& e : folded) if (e != proof.final_poly) throw std::runtime_error("FRI final poly is not constant");
  ch.observe_ext(proof.final_poly);
  proof.pow_witness = ch.grind(cfg.pow_bits);

  // ---- query phase -----------------------------------------------------------------------
  proof.query_proofs.clear();
  for (unsigned q = 0; q < cfg.num_queries; q++) {
    size_t index = ch.sample_bits(log_gmax);
    QueryProof qp;
    for (auto& r : rounds) {
      unsigned bits_reduced = log_gmax - r.data->tree.log_max_height();
      BatchOpening bo;
      r.data->tree.open_batch(index >> bits_reduced, bo.opened_values, bo.opening_proof);
      qp.input_proof.push_back(std::move(bo));
    }
    for (size_t i = 0; i < layer_trees.size(); i++) {
      size_t idx_i = index >> i, pair = idx_i >> 1;
      std::vector<std::vector<F>> rows;
      CommitPhaseStep st;
      layer_trees[i]->open_batch(pair, rows, st.opening_proof);
      const F* sib = rows[0].data() + 4 * ((idx_i ^ 1) & 1);
      st.sibling_value = E(sib[0], sib[1], sib[2], sib[3]);
      qp.commit_phase_openings.push_back(std::move(st));
    }
    proof.query_proofs.push_back(std::move(qp));
  }
}

// Verifier, restated from crates/recursion/circuit/src/fri.rs:34-361 (written from the circuit,
// independently of pcs_open above).
struct VerifyMat { Domain domain; std::vector<std::pair<E, std::vector<E>>> points_and_values; };
struct VerifyRound { Digest commit; std::vector<VerifyMat> mats; };

static inline std::string pcs_verify(const std::vector<VerifyRound>& rounds, const FriProof& proof,
                                     Challenger& ch, const FriConfig& cfg) {
  const E alpha = ch.sample_ext();
  std::vector<E> betas;
  for (auto& c : proof.commit_phase_commits) { ch.observe_digest(c); betas.push_back(ch.sample_ext()); }
  ch.observe_ext(proof.final_poly);
  if (proof.query_proofs.size() != cfg.num_queries) return "wrong number of queries";
  if (!ch.check_witness(cfg.pow_bits, proof.pow_witness)) return "invalid proof-of-work witness";
  const unsigned log_gmax = (unsigned)proof.commit_phase_commits.size() + cfg.log_blowup;
  for (auto& qp : proof.query_proofs) {
    const size_t index = ch.sample_bits(log_gmax);
    if (qp.input_proof.size() != rounds.size()) return "input proof round count";
    size_t log_height_pow[32] = {0};
    E ro[32];
    for (size_t ri = 0; ri < rounds.size(); ri++) {
      const VerifyRound& round = rounds[ri];
      const BatchOpening& bo = qp.input_proof[ri];
      std::vector<size_t> dims;
      for (auto& m : round.mats) dims.push_back((size_t)1 << (m.domain.log_n + cfg.log_blowup));
      if (bo.opened_values.size() != dims.size()) return "opened values count";
      size_t bmax = *std::max_element(dims.begin(), dims.end());
      unsigned lbmax = log2_strict(bmax);
      unsigned bits_reduced = log_gmax - lbmax;
      if (!verify_batch(round.commit, dims, index >> bits_reduced, lbmax, bo.opened_values, bo.opening_proof))
        return "input merkle proof rejected (round " + std::to_string(ri) + ")";
      for (size_t mi = 0; mi < round.mats.size(); mi++) {
        const VerifyMat& mat = round.mats[mi];
        unsigned lh = mat.domain.log_n + cfg.log_blowup;
        unsigned br = log_gmax - lh;
        size_t ridx = (index >> br) & (((size_t)1 << lh) - 1);
        F x = F(GENERATOR) * fpow(two_adic_generator(lh), bitrev(ridx, lh));
        for (auto& pv : mat.points_and_values) {
          if (pv.second.size() != bo.opened_values[mi].size()) return "opened width mismatch";
          E acc;
          for (size_t j = 0; j < pv.second.size(); j++) {
            acc += epow(alpha, log_height_pow[lh]) * (pv.second[j] - bo.opened_values[mi][j]);
            log_height_pow[lh]++;
          }
          ro[lh] += acc * einv(pv.first - x);
        }
      }
    }
    if (!ro[cfg.log_blowup].is_zero()) return "reduced opening at log_blowup is non-zero";
    // verify_query
    if (qp.commit_phase_openings.size() != proof.commit_phase_commits.size()) return "commit phase openings count";
    E folded = ro[log_gmax];
    F x = fpow(two_adic_generator(log_gmax), bitrev(index, log_gmax));
    for (size_t off = 0; off < proof.commit_phase_commits.size(); off++) {
      unsigned lfh = log_gmax - 1 - (unsigned)off;
      bool bit = (index >> off) & 1;
      size_t pair = index >> (off + 1);
      const CommitPhaseStep& st = qp.commit_phase_openings[off];
      E e0 = bit ? st.sibling_value : folded, e1 = bit ? folded : st.sibling_value;
      std::vector<F> leaf;
      for (int c = 0; c < 4; c++) leaf.push_back(e0.c[c]);
      for (int c = 0; c < 4; c++) leaf.push_back(e1.c[c]);
      if (!verify_batch(proof.commit_phase_commits[off], {(size_t)1 << lfh}, pair, lfh, {leaf}, st.opening_proof))
        return "commit phase merkle proof rejected (layer " + std::to_string(off) + ")";
      F xs0 = bit ? -x : x, xs1 = bit ? x : -x;
      folded = e0 + (betas[off] - xs0) * ((e1 - e0) * finv(xs1 - xs0));
      folded += betas[off] * betas[off] * ro[lfh];
      x = x * x;
    }
    if (folded != proof.final_poly) return "final poly mismatch";
  }
  return "";
}

}  // namespace zko

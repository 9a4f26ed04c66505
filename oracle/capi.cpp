// ORACLE (test infrastructure, not product code).  C entry points over the CPU restatement so
// that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can drive it with ctypes.
// All field elements cross this boundary in CANONICAL form.  Nothing under ziren_b200/ links,
// imports or calls this library.
#include "prover.h"
#include "tracegen.h"
#include "tracegen_global.h"
#include "tracegen_keccak.h"
#include <map>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace zko;

static thread_local std::string g_err;
static int fail(const std::exception& e) { g_err = e.what(); return 1; }

static Matrix mat_from(const u32* p, size_t h, size_t w) {
  Matrix m(h, w);
  for (size_t i = 0; i < h * w; i++) m.v[i] = F(p[i]);
  return m;
}
static E e_from(const u32* p) { return E(F(p[0]), F(p[1]), F(p[2]), F(p[3])); }
static void e_to(const E& e, u32* p) { for (int i = 0; i < 4; i++) p[i] = e.c[i].v; }

extern "C" {

const char* zko_last_error() { return g_err.c_str(); }
int zko_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void zko_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void zko_poseidon2_permute(u32* s) {
  F st[16];
  for (int i = 0; i < 16; i++) st[i] = F(s[i]);
  poseidon2_permute(st);
  for (int i = 0; i < 16; i++) s[i] = st[i].v;
}
void zko_poseidon2_permute_batch(u32* s, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < n; k++) zko_poseidon2_permute(s + 16 * k);
}
void zko_hash(const u32* in, size_t n, u32* out) {
  std::vector<F> v(n);
  for (size_t i = 0; i < n; i++) v[i] = F(in[i]);
  Digest d = sponge_hash(v.data(), n);
  for (int i = 0; i < 8; i++) out[i] = d[i].v;
}
void zko_compress(const u32* l, const u32* r, u32* out) {
  Digest a, b;
  for (int i = 0; i < 8; i++) { a[i] = F(l[i]); b[i] = F(r[i]); }
  Digest d = compress2(a, b);
  for (int i = 0; i < 8; i++) out[i] = d[i].v;
}
void zko_ef_mul(const u32* a, const u32* b, u32* out) { e_to(e_from(a) * e_from(b), out); }
void zko_ef_inv(const u32* a, u32* out) { e_to(einv(e_from(a)), out); }
u32 zko_two_adic_generator(unsigned bits) { return two_adic_generator(bits).v; }

void zko_dft(u32* data, size_t h, size_t w, int inverse) {
  Matrix m = mat_from(data, h, w);
  dft_rows(m, inverse != 0);
  for (size_t i = 0; i < h * w; i++) data[i] = m.v[i].v;
}
// out: (h << added_bits) x w, rows bit-reversed (the committed layout)
void zko_coset_lde(const u32* in, size_t h, size_t w, unsigned added_bits, u32 shift, u32* out) {
  Matrix r = coset_lde_bitrev(mat_from(in, h, w), added_bits, F(shift));
  for (size_t i = 0; i < r.v.size(); i++) out[i] = r.v[i].v;
}
// Merkle root over matrices as given (no LDE): MerkleTreeMmcs::commit
void zko_mmcs_root(int nmats, const u32* const* mats, const size_t* heights, const size_t* widths, u32* root) {
  MerkleTree t;
  for (int i = 0; i < nmats; i++) t.mats.push_back(mat_from(mats[i], heights[i], widths[i]));
  t.build();
  for (int i = 0; i < 8; i++) root[i] = t.root[i].v;
}
// TwoAdicFriPcs::commit root: LDE (shift GENERATOR/domain_shift) + Merkle
void zko_pcs_commit_root(int nmats, const u32* const* mats, const size_t* heights, const size_t* widths,
                         const u32* domain_shifts, unsigned log_blowup, u32* root) {
  FriConfig cfg; cfg.log_blowup = log_blowup;
  std::vector<std::pair<Domain, Matrix>> in;
  for (int i = 0; i < nmats; i++)
    in.push_back({Domain{log2_strict(heights[i]), F(domain_shifts ? domain_shifts[i] : 1)}, mat_from(mats[i], heights[i], widths[i])});
  auto cd = pcs_commit(in, cfg);
  for (int i = 0; i < 8; i++) root[i] = cd->commit()[i].v;
}
// one FRI fold step on m EF values (bit-reversed order), optional add of beta^2 * ro_next
void zko_fri_fold(const u32* in, size_t m, const u32* beta, const u32* ro_next, u32* out) {
  unsigned lm = log2_strict(m);
  F g = two_adic_generator(lm);
  E b = e_from(beta), b2 = b * b;
  for (size_t i = 0; i < m / 2; i++) {
    E r = fold_pair(e_from(in + 8 * i), e_from(in + 8 * i + 4), fpow(g, bitrev(2 * i, lm)), b);
    if (ro_next) r += b2 * e_from(ro_next + 4 * i);
    e_to(r, out + 4 * i);
  }
}
// challenger script: ops[i] = 0 observe(vals[i]) | 1 sample -> out | 2 sample_bits(vals[i]) -> out
void zko_challenger_script(u32* state34, const u32* ops, const u32* vals, size_t n, u32* out) {
  Challenger ch; ch.from_words(state34);
  size_t o = 0;
  for (size_t i = 0; i < n; i++) {
    if (ops[i] == 0) ch.observe(F(vals[i]));
    else if (ops[i] == 1) out[o++] = ch.sample().v;
    else out[o++] = ch.sample_bits(vals[i]);
  }
  ch.to_words(state34);
}
u32 zko_grind(u32* state34, unsigned bits) {
  Challenger ch; ch.from_words(state34);
  F w = ch.grind(bits);
  ch.to_words(state34);
  return w.v;
}

// ---- machine level ------------------------------------------------------------------------------
void* zko_machine_new(const u32* desc, size_t n) {
  try { return new Machine(parse_machine(desc, n)); } catch (const std::exception& e) { fail(e); return nullptr; }
}
void zko_machine_free(void* m) { delete (Machine*)m; }
int zko_chip_info(void* m_, const char* name, u32* out /* perm_width_ef, num_constraints, lqd */) {
  const Chip* c = ((Machine*)m_)->find(name);
  if (!c) return 1;
  out[0] = (u32)c->perm_width_ef(); out[1] = (u32)c->num_constraints(); out[2] = c->log_quotient_degree;
  return 0;
}

void* zko_setup(void* m_, int n, const char* const* names, const u32* const* ptrs, const size_t* heights,
                const size_t* widths, u32 pc_start, const u32* init_gsum, u32* commit_out) {
  try {
    std::vector<std::pair<std::string, Matrix>> prep;
    for (int i = 0; i < n; i++) prep.push_back({names[i], mat_from(ptrs[i], heights[i], widths[i])});
    F gs[14];
    for (int i = 0; i < 14; i++) gs[i] = F(init_gsum ? init_gsum[i] : SEPTIC_DIGEST_ZERO[i]);
    auto pk = setup(*(Machine*)m_, std::move(prep), F(pc_start), gs);
    if (commit_out) for (int i = 0; i < 8; i++) commit_out[i] = pk->commit[i].v;
    return pk.release();
  } catch (const std::exception& e) { fail(e); return nullptr; }
}
void zko_pk_free(void* pk) { delete (ProvingKey*)pk; }
// the challenger every shard proof starts from: fresh challenger after pk.observe_into
void zko_pk_initial_challenger(void* pk, u32* state34) {
  Challenger ch;
  ((ProvingKey*)pk)->observe_into(ch);
  ch.to_words(state34);
}

// commit + open of one shard.  challenger34 is in/out (a clone of the post-observe_into state).
int zko_prove_shard(void* m_, void* pk_, int n, const char* const* names, const u32* const* ptrs,
                    const size_t* heights, const size_t* widths, const u32* pv, size_t npv, u32* challenger34,
                    u32** out_words, size_t* out_len) {
  try {
    const Machine& m = *(Machine*)m_;
    std::vector<std::pair<std::string, Matrix>> tr;
    for (int i = 0; i < n; i++) tr.push_back({names[i], mat_from(ptrs[i], heights[i], widths[i])});
    std::vector<F> pvs(npv);
    for (size_t i = 0; i < npv; i++) pvs[i] = F(pv[i]);
    auto sd = commit(m, std::move(tr), pvs);
    Challenger ch; ch.from_words(challenger34);
    auto proof = open(m, *(ProvingKey*)pk_, *sd, ch);
    ch.to_words(challenger34);
    std::vector<u32> w = serialize_proof(*proof);
    *out_words = (u32*)malloc(w.size() * 4);
    memcpy(*out_words, w.data(), w.size() * 4);
    *out_len = w.size();
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
// main commitment only (MachineProver::commit)
int zko_commit_shard(void* m_, int n, const char* const* names, const u32* const* ptrs, const size_t* heights,
                     const size_t* widths, u32* commit_out) {
  try {
    std::vector<std::pair<std::string, Matrix>> tr;
    for (int i = 0; i < n; i++) tr.push_back({names[i], mat_from(ptrs[i], heights[i], widths[i])});
    auto sd = commit(*(Machine*)m_, std::move(tr), {});
    for (int i = 0; i < 8; i++) commit_out[i] = sd->main_data->commit()[i].v;
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
void zko_free(void* p) { free(p); }

// returns 0 when the proof verifies; otherwise 1 and zko_last_error() says why
int zko_verify_shard(void* m_, void* pk_, const u32* proof_words, size_t n, const u32* challenger34) {
  try {
    ShardProof p = parse_proof(proof_words, n);
    VerifyingKey vk = vk_from_pk(*(ProvingKey*)pk_);
    Challenger ch; ch.from_words(challenger34);
    std::string err = verify_shard(*(Machine*)m_, vk, ch, p);
    if (!err.empty()) { g_err = err; return 1; }
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

// ---- stage hooks for kernel-level parity tests ----------------------------------------------------
int zko_permutation_trace(void* m_, const char* chip, const u32* prep, const u32* main, size_t height,
                          const u32* alpha, const u32* beta, u32* out, u32* local_sum) {
  try {
    const Chip* c = ((Machine*)m_)->find(chip);
    if (!c) throw std::runtime_error("unknown chip");
    Matrix pm, mm = mat_from(main, height, c->main_width);
    if (prep) pm = mat_from(prep, height, c->prep_width);
    E ls;
    Matrix r = generate_permutation_trace(*c, prep ? &pm : nullptr, mm, e_from(alpha), e_from(beta), ls);
    for (size_t i = 0; i < r.v.size(); i++) out[i] = r.v[i].v;
    e_to(ls, local_sum);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
// The multiplicity columns of a receive-only table from the rows of its senders (the checker of csrc/derive.cuh): what the
// reference accumulates on the host as record.byte_lookups -> ByteChip::generate_trace (crates/core/machine/src/bytes/trace.rs:
// 46-67) and as instruction counts -> ProgramChip::generate_trace (program/mod.rs:115-158), stated through the lookups of the
// machine description - a std::map from (kind, tuple) to the first (receive, row) of the receiver that holds it.
// Tables row-major canonical; senders: n_senders x {chip name, prep or null, main, height}; out height x main_width canonical.
int zko_derive_multiplicities(void* m_, const char* receiver, const u32* receiver_prep, size_t receiver_height, int n_senders,
                              const char* const* sender_names, const u32* const* sender_prep, const u32* const* sender_main,
                              const size_t* sender_heights, u32* out, unsigned long long* n_lookups) {
  try {
    const Machine& m = *(Machine*)m_;
    const Chip* r = m.find(receiver);
    if (!r) throw std::runtime_error("unknown chip");
    std::map<std::vector<u32>, std::pair<size_t, u32>> where;      // (kind, values...) -> (row, multiplicity column)
    std::vector<u32> kinds;
    for (auto& l : r->receives) {
      if (l.scope != 0) continue;
      bool prep_only = true;
      for (auto& v : l.values) for (auto& t : v.terms) prep_only = prep_only && !t.is_main;
      if (!prep_only || !l.mult.constant.is_zero() || l.mult.terms.size() != 1 || !l.mult.terms[0].is_main || l.mult.terms[0].w != F::one()) continue;
      kinds.push_back(l.kind);
      for (size_t row = 0; row < receiver_height; row++) {
        std::vector<u32> key{l.kind};
        for (auto& v : l.values) key.push_back(v.apply<F, F>((const F*)receiver_prep + row * r->prep_width, (const F*)nullptr).v);
        where.emplace(key, std::make_pair(row, l.mult.terms[0].col));      // the first (receive, row) keeps a repeated tuple
      }
    }
    if (kinds.empty()) throw std::runtime_error("no receive of preprocessed columns with a main column as its multiplicity");
    std::vector<u64> counts(receiver_height * r->main_width, 0);
    unsigned long long n = 0;
    for (int i = 0; i < n_senders; i++) {
      const Chip* c = m.find(sender_names[i]);
      if (!c) throw std::runtime_error("unknown chip");
      for (auto& l : c->sends) {
        if (l.scope != 0 || std::find(kinds.begin(), kinds.end(), l.kind) == kinds.end()) continue;
        for (size_t row = 0; row < sender_heights[i]; row++) {
          const F* pr = sender_prep[i] ? (const F*)sender_prep[i] + row * c->prep_width : nullptr;
          const F* mr = (const F*)sender_main[i] + row * c->main_width;
          const F mult = l.mult.apply<F, F>(pr, mr);
          if (mult.is_zero()) continue;
          std::vector<u32> key{l.kind};
          for (auto& v : l.values) key.push_back(v.apply<F, F>(pr, mr).v);
          auto it = where.find(key);
          if (it == where.end()) throw std::runtime_error("a lookup of " + c->name + " is in no row of " + r->name);
          counts[it->second.first * r->main_width + it->second.second] += mult.v;
          n++;
        }
      }
    }
    for (size_t i = 0; i < counts.size(); i++) out[i] = (u32)(counts[i] % P);
    if (n_lookups) *n_lookups = n;
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
// LDE inputs are the committed (bit-reversed) matrices of height n << log_blowup; out is Q x 4 natural order
int zko_quotient_values(void* m_, const char* chip, unsigned log_n, const u32* prep_lde, const u32* main_lde,
                        const u32* perm_lde, const u32* perm_alpha, const u32* perm_beta, const u32* local_sum,
                        const u32* global_sum, const u32* alpha, const u32* pub, size_t npub, u32* out) {
  try {
    const Machine& m = *(Machine*)m_;
    const Chip* c = m.find(chip);
    if (!c) throw std::runtime_error("unknown chip");
    size_t H = (size_t)1 << (log_n + m.cfg.log_blowup);
    Matrix pl, ml = mat_from(main_lde, H, c->main_width), el = mat_from(perm_lde, H, 4 * c->perm_width_ef());
    if (prep_lde) pl = mat_from(prep_lde, H, c->prep_width);
    F gs[14];
    for (int i = 0; i < 14; i++) gs[i] = F(global_sum[i]);
    std::vector<F> pubv(npub + 1);
    for (size_t i = 0; i < npub; i++) pubv[i] = F(pub[i]);
    std::vector<E> q = quotient_values(*c, log_n, prep_lde ? &pl : nullptr, ml, el, e_from(perm_alpha), e_from(perm_beta),
                                       e_from(local_sum), gs, e_from(alpha), pubv.data());
    for (size_t i = 0; i < q.size(); i++) e_to(q[i], out + 4 * i);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}

// trace generation of the core ALU chips (tracegen.h): events n x 7 words {pc, next_pc, opcode, hi, a, b, c},
// out height x width row-major canonical
int zko_alu_width(int chip) { return chip >= 0 && chip < T_NCHIPS ? ALU_WIDTHS[chip] : -1; }
int zko_alu_trace(int chip, const u32* ev, size_t n, size_t height, u32* out) {
  try {
    std::vector<AluEvent> e(n);
    for (size_t i = 0; i < n; i++)
      e[i] = AluEvent{ev[7 * i], ev[7 * i + 1], ev[7 * i + 2], ev[7 * i + 3], ev[7 * i + 4], ev[7 * i + 5], ev[7 * i + 6]};
    // the #[repr(u8)] opcode is the low byte of its word: word 2 of an AluEvent, word 3 of a Branch/JumpEvent
    for (auto& x : e) { if (chip == T_BRANCH || chip == T_JUMP) x.hi &= 0xff; else x.opcode &= 0xff; }
    alu_trace(chip, e.data(), n, height, out);
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return 1;
  }
}

// MulChip rows (tracegen.h): events n x 16 words (CompAluEvent), out height x 58 row-major canonical
int zko_mul_trace(const u32* ev, size_t n, size_t height, u32* out) {
  try {
    mul_trace(ev, n, height, out);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
// MemoryInstrs rows (tracegen.h): events n x 16 words (MemInstrEvent), out height x 79 row-major canonical
int zko_mem_instr_trace(const u32* ev, size_t n, size_t height, u32* out) {
  try {
    mem_instr_trace(ev, n, height, out);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
// MemoryLocal rows (tracegen.h): events n x 7 words (MemoryLocalEvent), four per row, out height x 56 row-major canonical
int zko_memory_local_trace(const u32* ev, size_t n, size_t height, u32* out) {
  try {
    memory_local_trace(ev, n, height, out);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
// Cpu rows (tracegen.h): events n x 28 words (zkb200_cpu_event), out height x 67 row-major canonical
int zko_cpu_trace(const u32* ev, size_t n, size_t height, u32* out) {
  try {
    cpu_trace(ev, n, height, out);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
// MiscInstrs rows (tracegen.h): events n x 15 words (MiscEvent), out height x 72 row-major canonical
int zko_misc_trace(const u32* ev, size_t n, size_t height, u32* out) {
  try {
    misc_trace(ev, n, height, out);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
// DivRem / SyscallCore / SyscallPrecompile / SyscallInstrs rows (tracegen.h chip_trace): events n x chip_event_words, out
// height x chip_trace_width row-major canonical
int zko_chip_trace_width(const char* chip) { return chip_trace_width(chip); }
int zko_chip_event_words(const char* chip) { return chip_event_words(chip); }
int zko_chip_trace(const char* chip, const u32* ev, size_t n, size_t height, u32* out) {
  try {
    chip_trace(chip, ev, n, height, out);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
// MemoryGlobalInit / MemoryGlobalFinalize rows (tracegen.h memory_global_trace): address-sorted events n x 4 words and the
// public values' previous address, out height x 111 row-major canonical
int zko_memory_global_trace(const u32* ev, size_t n, u32 previous_addr, size_t height, u32* out) {
  try {
    memory_global_trace(ev, n, previous_addr, height, out);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
// Global rows (tracegen_global.h): events n x 8 words (GlobalLookupEvent), out height x 99 row-major canonical
int zko_global_trace(const u32* ev, size_t n, size_t height, u32* out) {
  try {
    global_trace(ev, n, height, out);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
// septic extension / curve primitives of tracegen_global.h, canonical words: op 0 a * b, 1 1 / a, 2 sqrt(a) (returns 1 when a is
// not a square), 3 frobenius(a), 4 double_frobenius(a), 5 curve_formula(a) (7 words each); 6 complete addition of two points
// (a, b, out: 14 words each, the point at infinity as zeros)
int zko_septic_op(int op, const u32* a, const u32* b, u32* out) {
  try {
    Septic x, y, r;
    for (int i = 0; i < 7; i++) { x.c[i] = F(a[i]); if (b) y.c[i] = F(b[i]); }
    if (op == 6) {
      CurvePoint p = curve_point_from_words(a), q = curve_point_from_words(b);
      p.infinity = p.x.is_zero() && p.y.is_zero(); q.infinity = q.x.is_zero() && q.y.is_zero();
      const CurvePoint s = curve_add_complete(p, q);
      for (int i = 0; i < 7; i++) { out[i] = s.infinity ? 0 : s.x.c[i].v; out[7 + i] = s.infinity ? 0 : s.y.c[i].v; }
      return 0;
    }
    switch (op) {
      case 0: r = x * y; break;
      case 1: r = septic_inv(x); break;
      case 2: if (!septic_sqrt(x, r)) return 1; break;
      case 3: r = septic_frobenius(x); break;
      case 4: r = septic_double_frobenius(x); break;
      case 5: r = curve_formula(x); break;
      default: throw std::runtime_error("oracle: unknown septic op");
    }
    for (int i = 0; i < 7; i++) out[i] = r.c[i].v;
    return 0;
  } catch (const std::exception& e) { fail(e); return -1; }
}
// trace generation of the KeccakSponge chip (tracegen_keccak.h): n_blocks records of KS_REC_WORDS words,
// out height x 3531 row-major canonical
int zko_keccak_sponge_width() { return KS_WIDTH; }
int zko_keccak_sponge_trace(const u32* recs, size_t n_blocks, size_t height, u32* out) {
  try {
    keccak_sponge_trace(recs, n_blocks, height, out);
    return 0;
  } catch (const std::exception& e) { return fail(e); }
}
// MemoryAccessCols::populate_access alone (9 canonical words), for the comparison with the reference's memory.hpp
void zko_mem_access(u32 value, u32 shard, u32 ts, u32 prev_shard, u32 prev_ts, u32* out9) { ks_mem_access(out9, value, shard, ts, prev_shard, prev_ts); }

}  // extern "C"

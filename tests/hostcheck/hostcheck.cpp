// Host build of the product's arithmetic headers (ziren_b200/csrc/kb31.cuh, poseidon2.cuh): the
// device kernels run exactly these expressions, so the CPU suite can pin them against the golden
// vectors and the oracle without a GPU.  Test infrastructure only (built by tests/test_host_logic.py).
#include "poseidon2.cuh"
#include "tracegen.cuh"
#include "tracegen_keccak.cuh"
#include "tracegen_global.cuh"
#include "derive.cuh"
#include "lane_pool.h"
#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

namespace zkb {
const P2Consts& p2_host_consts() {
  static const P2Consts c = p2_make_consts();
  return c;
}
}  // namespace zkb
using namespace zkb;

// the CTA of the row kernel (csrc/tracegen.cuh AluCta, the three phases csrc/tracegen.cu alu_rows_kernel runs between its
// barriers) walked thread by thread on the host: the same load / fill / store index arithmetic as on the GPU, shared memory
// as two arrays per CTA.  Poisoned shared memory and output catch reads of words the load phase did not bring in and rows
// the store phase does not write.
template <int CHIP>
static void walk_ctas(const uint32_t* ev, size_t n, size_t height, uint32_t* out, int col_major, const u32* inv255) {
  using C = AluCta<CHIP>;
  std::vector<u32> ev_s((size_t)C::R * C::RW), tile((size_t)C::R * C::WP);
  const size_t ctas = (height + C::R - 1) / C::R;
  for (size_t cta = 0; cta < ctas; cta++) {
    std::fill(ev_s.begin(), ev_s.end(), 0xDEADBEEFu);
    std::fill(tile.begin(), tile.end(), 0xDEADBEEFu);
    for (u32 t = 0; t < (u32)C::R; t++) C::load(t, cta, ev, n, ev_s.data());
    for (u32 t = 0; t < (u32)C::R; t++) C::fill(t, cta, n, ev_s.data(), tile.data(), inv255);
    for (u32 t = 0; t < (u32)C::R; t++) C::store(t, cta, height, tile.data(), out, col_major);
  }
}
template <int CHIP>
static int walk_dispatch(int chip, const uint32_t* ev, size_t n, size_t height, uint32_t* out, int col_major, const u32* inv255) {
  if (chip == CHIP) { walk_ctas<CHIP>(ev, n, height, out, col_major, inv255); return 0; }
  if constexpr (CHIP + 1 < ALU_NCHIPS) return walk_dispatch<CHIP + 1>(chip, ev, n, height, out, col_major, inv255);
  return 1;
}
extern "C" {
// canonical in / out, n states of 16 words
void hostcheck_permute(uint32_t* st, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Fp s[16];
    for (int j = 0; j < 16; j++) s[j] = fp_from_canonical(st[16 * i + j]);
    p2_permute_host(s);
    for (int j = 0; j < 16; j++) st[16 * i + j] = fp_to_canonical(s[j]);
  }
}
// op: 0 mul, 1 add, 2 sub, 3 inv(a), 4 to_monty(a), 5 halve(a); canonical in / out except op 4
void hostcheck_field(int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
  for (size_t i = 0; i < n; i++) {
    Fp x = fp_from_canonical(a[i]), y = fp_from_canonical(b ? b[i] : 0);
    switch (op) {
      case 0: out[i] = fp_to_canonical(x * y); break;
      case 1: out[i] = fp_to_canonical(x + y); break;
      case 2: out[i] = fp_to_canonical(x - y); break;
      case 3: out[i] = fp_to_canonical(fp_inv(x)); break;
      case 4: out[i] = x.v; break;
      case 5: out[i] = fp_to_canonical(fp_halve(x)); break;
    }
  }
}
// EF4 product, canonical coefficients
void hostcheck_ef_mul(const uint32_t* a, const uint32_t* b, uint32_t* out) {
  Ef x, y;
  for (int i = 0; i < 4; i++) { x.c[i] = fp_from_canonical(a[i]); y.c[i] = fp_from_canonical(b[i]); }
  Ef z = x * y;
  for (int i = 0; i < 4; i++) out[i] = fp_to_canonical(z.c[i]);
}
// lazy accumulation path used by the opening / quotient kernels: sum_i w_i * x_i (EF x base)
void hostcheck_efacc(const uint32_t* w, const uint32_t* x, size_t n, uint32_t* out) {
  EfAcc acc;
  acc.clear();
  for (size_t i = 0; i < n; i++) {
    Ef e;
    for (int j = 0; j < 4; j++) e.c[j] = fp_from_canonical(w[4 * i + j]);
    acc.add(e, fp_from_canonical(x[i]));
  }
  Ef r = acc.value();
  for (int j = 0; j < 4; j++) out[j] = fp_to_canonical(r.c[j]);
}
// the product's ALU row fillers (csrc/tracegen.cuh) on the host: events n x 7 words, out height x width
// row-major Montgomery, padding rows past the last event
int hostcheck_alu_rows(int chip, const uint32_t* ev, size_t n, size_t height, uint32_t* out) {
  if (chip < 0 || chip >= ALU_NCHIPS) return 1;
  const size_t epr = (size_t)alu_events_per_row(chip);
  if ((n + epr - 1) / epr > height) return 1;
  static u32 inv255[256];
  static bool init = false;
  if (!init) { alu_build_inv255(inv255); init = true; }
  const int w = alu_width(chip);
  for (size_t i = 0; i < height; i++) {
    if (i * epr < n) fill_alu_row(chip, ev + (size_t)alu_event_words(chip) * epr * i, out + i * w, inv255, (int)(n - i * epr < epr ? n - i * epr : epr));
    else fill_alu_padding(chip, out + i * w);
  }
  return 0;
}
int hostcheck_alu_width(int chip) { return alu_width(chip); }
int hostcheck_alu_rows_cta(int chip, const uint32_t* ev, size_t n, size_t height, uint32_t* out, int col_major) {
  if (chip < 0 || chip >= ALU_NCHIPS) return 1;
  const size_t epr = (size_t)alu_events_per_row(chip);
  if ((n + epr - 1) / epr > height) return 1;
  static u32 inv255[256];
  static bool init = false;
  if (!init) { alu_build_inv255(inv255); init = true; }
  return walk_dispatch<0>(chip, ev, n, height, out, col_major, inv255);
}
int hostcheck_alu_cta_rows(int chip) { return alu_cta_rows(chip); }
int hostcheck_alu_event_words(int chip) { return alu_event_words(chip); }
int hostcheck_alu_nchips() { return ALU_NCHIPS; }
// the Global chip's three steps (csrc/tracegen_global.cuh; csrc/tracegen.cu global_trace) walked on the host in the kernels'
// order: lift per event, the chunked scan level by level (totals, recursion, rescan: one call per GPU thread), finish per row
static const GlobalConsts& host_global_consts() {
  static const GlobalConsts k = [] { GlobalConsts c; global_build_consts(c); return c; }();
  return k;
}
static void host_global_scan(u32* pts, size_t n, const GlobalConsts& k) {
  if (n <= 1) return;
  const size_t chunks = (n + GLOBAL_SCAN_CHUNK - 1) / GLOBAL_SCAN_CHUNK;
  if (chunks == 1) { global_chunk_rescan(pts, n, 0, k, nullptr); return; }
  std::vector<u32> totals(chunks * GLOBAL_POINT_WORDS, 0xDEADBEEFu);
  for (size_t t = 0; t < chunks; t++) global_chunk_total(pts, n, t, k, totals.data());
  host_global_scan(totals.data(), chunks, k);
  for (size_t t = 0; t < chunks; t++) global_chunk_rescan(pts, n, t, k, totals.data());
}
int hostcheck_global_rows(const uint32_t* ev, size_t n, size_t height, uint32_t* out, int col_major) {
  if (n > height) return 1;
  const GlobalConsts& k = host_global_consts();
  const GlobalOut o{out, col_major ? (size_t)1 : (size_t)GLOBAL_WIDTH, col_major ? height : (size_t)1};
  std::vector<u32> points((n + 1) * GLOBAL_POINT_WORDS, 0xDEADBEEFu);
  curve_store(points.data(), k.start);
  for (size_t row = 0; row < n; row++) global_lift_row(ev + GLOBAL_EVENT_WORDS * row, row, k, o, points.data());
  host_global_scan(points.data(), n + 1, k);
  for (size_t row = 0; row < height; row++) global_finish_row(row, n, points.data(), k, o);
  return 0;
}
int hostcheck_global_width() { return GLOBAL_WIDTH; }
// septic primitives, canonical words in and out; ops as oracle/capi.cpp zko_septic_op
int hostcheck_septic_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  const GlobalConsts& k = host_global_consts();
  Sep x, y = sep_zero(), r;
  for (int i = 0; i < 7; i++) { x.c[i] = fp_from_canonical(a[i]); if (b) y.c[i] = fp_from_canonical(b[i]); }
  if (op == 6) {
    CurvePt p, q;
    for (int i = 0; i < 7; i++) {
      p.x.c[i] = fp_from_canonical(a[i]); p.y.c[i] = fp_from_canonical(a[7 + i]);
      q.x.c[i] = fp_from_canonical(b[i]); q.y.c[i] = fp_from_canonical(b[7 + i]);
    }
    const CurvePt s = curve_add(p, q, k);
    for (int i = 0; i < 7; i++) { out[i] = fp_to_canonical(s.x.c[i]); out[7 + i] = fp_to_canonical(s.y.c[i]); }
    return 0;
  }
  switch (op) {
    case 0: r = x * y; break;
    case 1: r = sep_inv(x, k); break;
    case 2: if (!sep_sqrt(x, k, r)) return 1; break;
    case 3: r = sep_frobenius(x, k); break;
    case 4: r = sep_double_frobenius(x, k); break;
    case 5: r = curve_formula(x); break;
    default: return -1;
  }
  for (int i = 0; i < 7; i++) out[i] = fp_to_canonical(r.c[i]);
  return 0;
}
// K7 (csrc/derive.cuh; csrc/derive.cu derive_multiplicities) walked on the host in the kernels' order: fill, one insert per
// (receive, row), one probe per (send, row), finish.  The machine tables arrive as the flat arrays MachineInfo::upload builds
// (DevLookup 5 words, DevVPC 3, DevTerm 2, Montgomery weights); tables column-major Montgomery.  stats: lookups counted, misses.
int hostcheck_derive(const uint32_t* lookups, const uint32_t* vpcs, const uint32_t* terms, const uint32_t* recv, uint32_t n_recv,
                     const uint32_t* receiver_prep, size_t receiver_height, uint32_t main_width, int n_sends, const uint32_t* send_lookup,
                     const uint32_t* send_table, const uint32_t* const* table_prep, const uint32_t* const* table_main,
                     const size_t* table_height, uint32_t* out, unsigned long long* stats) {
  static_assert(sizeof(DevLookup) == 20 && sizeof(DevVPC) == 12 && sizeof(DevTerm) == 8 && sizeof(DeriveReceive) == 8, "flat table layout");
  const DevLookup* L = reinterpret_cast<const DevLookup*>(lookups);
  const DevVPC* V = reinterpret_cast<const DevVPC*>(vpcs);
  const DevTerm* T = reinterpret_cast<const DevTerm*>(terms);
  const DeriveReceive* R = reinterpret_cast<const DeriveReceive*>(recv);
  const size_t entries = (size_t)n_recv * receiver_height;
  size_t cap = 1;
  while (cap < 2 * entries) cap <<= 1;
  std::vector<u32> slots(cap, DERIVE_EMPTY);
  std::vector<u64> counts((size_t)main_width * receiver_height, 0);
  const DeriveTable r{receiver_prep, nullptr, receiver_height};
  for (size_t i = 0; i < entries; i++) derive_insert((u32)(i / receiver_height), i % receiver_height, R, L, V, T, r, slots.data(), (u32)(cap - 1));
  stats[0] = stats[1] = 0;
  for (int k = 0; k < n_sends; k++) {
    const DeriveTable st{table_prep[send_table[k]], table_main[send_table[k]], table_height[send_table[k]]};
    for (size_t row = 0; row < st.height; row++) {
      const int rc = derive_probe(L[send_lookup[k]], row, st, R, L, V, T, r, slots.data(), (u32)(cap - 1), counts.data());
      if (rc) stats[rc - 1]++;
    }
  }
  for (size_t i = 0; i < counts.size(); i++) out[i] = derive_finish(counts[i]);
  return 0;
}
// the product's KeccakSponge row filler (csrc/tracegen_keccak.cuh) on the host: n_blocks records of 384 words,
// out height x 3531 row-major Montgomery, padding rows past the last block
struct HostRowStore { uint32_t* r; void operator()(int col, u32 v) { r[col] = v; } };
int hostcheck_keccak_rows(const uint32_t* recs, size_t n_blocks, size_t height, uint32_t* out) {
  if (n_blocks * KS_ROUNDS > height) return 1;
  for (size_t i = 0; i < height; i++) {
    const size_t b = i / KS_ROUNDS;
    HostRowStore st{out + i * KS_WIDTH};
    ks_fill_row(b < n_blocks ? recs + b * KS_REC_WORDS : nullptr, (u32)(i % KS_ROUNDS), st);
  }
  return 0;
}
int hostcheck_keccak_width() { return KS_WIDTH; }
// stress of the lane pool (csrc/lane_pool.h): `threads` host threads take and release lanes `iters` times.
// Returns 0 when no lane ever had two holders and never more than `active` lanes were held at once.
int hostcheck_lane_pool(int threads, int iters, int active) {
  zkb::LanePool<4> pool;
  std::atomic<int> holders[4];
  std::atomic<long> uses[4];
  for (int i = 0; i < 4; i++) { holders[i] = 0; uses[i] = 0; }
  std::atomic<int> held{0}, bad{0};
  std::vector<std::thread> ts;
  for (int t = 0; t < threads; t++)
    ts.emplace_back([&, t] {
      unsigned x = 12345u + 977u * (unsigned)t;
      for (int i = 0; i < iters; i++) {
        const int l = pool.acquire(active);
        if (l < 0 || l >= active) bad |= 1;
        if (holders[l].fetch_add(1) != 0) bad |= 2;
        if (held.fetch_add(1) >= active) bad |= 4;
        uses[l]++;
        x = x * 1664525u + 1013904223u;
        for (volatile unsigned spin = 0; spin < (x >> 24); spin++) {}
        held.fetch_sub(1);
        holders[l].fetch_sub(1);
        pool.release(l);
      }
    });
  for (auto& t : ts) t.join();
  if (uses[0] == 0) bad |= 8;      // lane 0 is always the first choice; higher lanes only see use under contention
  return bad.load();
}
}

// K7: LogUp multiplicities of a table that only RECEIVES (Byte, Program, range tables), derived on the device from the rows
// of the tables that send to it - the host-side counterpart in the reference is every chip's generate_dependencies feeding
// record.byte_lookups / the Cpu events feeding the Program table (crates/core/machine/src/bytes/trace.rs:46-67,
// program/mod.rs:115-158): a histogram of the lookups the AIRs send.  Here the AIR itself says which tuples a row sends (the
// descriptor's sends, the same VirtualPairCols K5 turns into the permutation trace), so one generic pass over resident rows -
// uploaded or generated on the device - replaces the per-chip byte-lookup bookkeeping:
//   build   every (receive, row) of the receiver whose tuple is made of preprocessed columns only goes into an open-addressing
//           hash table of entry ids; tuples are never stored, an entry's tuple is recomputed from the preprocessed table
//   probe   every (send, row) of a sender with a non-zero multiplicity is looked up (kind, length and every value compared)
//           and its multiplicity added to the receiver's multiplicity column of the matching row; a tuple that is in no table
//           is counted as a miss and the call fails
// Host/device code: tests/hostcheck walks the same functions thread by thread.
#pragma once
#include "machine_dev.h"

namespace zkb {

constexpr u32 DERIVE_EMPTY = 0xffffffffu, DERIVE_MAX_VALUES = 16;

struct DeriveTable {               // one resident table, column-major Montgomery
  const u32* prep; const u32* main_trace; size_t height;
};
// VirtualPairCol::apply (crates/stark/src/air/...): constant + sum of weight * column
KB_HD Fp derive_eval_vpc(const DevVPC& v, const DevTerm* terms, const DeriveTable& t, size_t row) {
  Fp acc = fp_raw(v.constant);
  for (u32 i = v.term_begin; i < v.term_end; i++) {
    const u32 col = terms[i].col & 0x7fffffffu;
    const u32 x = (terms[i].col >> 31) ? t.main_trace[(size_t)col * t.height + row] : t.prep[(size_t)col * t.height + row];
    acc += fp_raw(terms[i].w) * fp_raw(x);
  }
  return acc;
}
// the tuple of lookup `l` at `row`, canonical; returns its length
KB_HD u32 derive_tuple(const DevLookup& l, const DevVPC* vpcs, const DevTerm* terms, const DeriveTable& t, size_t row, u32* out) {
  const u32 n = l.value_end - l.value_begin;
  for (u32 j = 0; j < n; j++) out[j] = fp_to_canonical(derive_eval_vpc(vpcs[l.value_begin + j], terms, t, row));
  return n;
}
KB_HD u64 derive_hash(u32 kind, const u32* v, u32 n) {
  u64 h = 0x9e3779b97f4a7c15ull ^ ((u64)kind << 32 | n);
  for (u32 j = 0; j < n; j++) {
    h ^= v[j];
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 29;
  }
  return h ^ (h >> 32);
}
#if defined(__CUDA_ARCH__)
#define ZKB_ATOMIC_CAS_U32(p, cmp, val) atomicCAS((p), (cmp), (val))
#define ZKB_ATOMIC_ADD_U64(p, val) atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)(val))
#else
static inline u32 zkb_host_cas(u32* p, u32 cmp, u32 val) { const u32 old = *p; if (old == cmp) *p = val; return old; }
#define ZKB_ATOMIC_CAS_U32(p, cmp, val) zkb_host_cas((p), (cmp), (val))
#define ZKB_ATOMIC_ADD_U64(p, val) (*(p) += (val))
#endif

// the receiver's eligible receives: lookup index (machine-wide), the main column that holds its multiplicity
struct DeriveReceive { u32 lookup; u32 mult_col; };

// build: entry id = receive_slot * height + row
KB_HD void derive_insert(u32 receive_slot, size_t row, const DeriveReceive* recv, const DevLookup* lookups, const DevVPC* vpcs,
                         const DevTerm* terms, const DeriveTable& r, u32* slots, u32 cap_mask) {
  u32 v[DERIVE_MAX_VALUES];
  const DevLookup& l = lookups[recv[receive_slot].lookup];
  const u32 n = derive_tuple(l, vpcs, terms, r, row, v);
  const u32 id = receive_slot * (u32)r.height + (u32)row;
  u32 s = (u32)derive_hash(l.kind, v, n) & cap_mask;
  while (ZKB_ATOMIC_CAS_U32(slots + s, DERIVE_EMPTY, id) != DERIVE_EMPTY) s = (s + 1) & cap_mask;
}
// probe: one send of one sender row.  Among equal tuples of the receiver the entry with the smallest id takes the
// multiplicity (the reference's fixed tables hold every tuple once).  Returns 0 no lookup (multiplicity zero), 1 counted, 2 miss.
KB_HD int derive_probe(const DevLookup& send, size_t row, const DeriveTable& s_tab, const DeriveReceive* recv, const DevLookup* lookups,
                       const DevVPC* vpcs, const DevTerm* terms, const DeriveTable& r, const u32* slots, u32 cap_mask, u64* counts) {
  const u32 mult = fp_to_canonical(derive_eval_vpc(vpcs[send.mult_vpc], terms, s_tab, row));
  if (!mult) return 0;
  u32 v[DERIVE_MAX_VALUES], w[DERIVE_MAX_VALUES];
  const u32 n = derive_tuple(send, vpcs, terms, s_tab, row, v);
  u32 best = DERIVE_EMPTY;
  for (u32 s = (u32)derive_hash(send.kind, v, n) & cap_mask;; s = (s + 1) & cap_mask) {
    const u32 id = slots[s];
    if (id == DERIVE_EMPTY) break;
    if (id >= best) continue;
    const u32 slot = id / (u32)r.height;
    const size_t rrow = id % (u32)r.height;
    const DevLookup& l = lookups[recv[slot].lookup];
    if (l.kind != send.kind || l.value_end - l.value_begin != n) continue;
    derive_tuple(l, vpcs, terms, r, rrow, w);
    bool same = true;
    for (u32 j = 0; j < n; j++) same = same && v[j] == w[j];
    if (same) best = id;
  }
  if (best == DERIVE_EMPTY) return 2;
  ZKB_ATOMIC_ADD_U64(counts + (size_t)recv[best / (u32)r.height].mult_col * r.height + best % (u32)r.height, mult);
  return 1;
}
// counts (sums of canonical multiplicities) -> the receiver's main trace, Montgomery
KB_HD u32 derive_finish(u64 count) { return fp_from_canonical((u32)(count % KB_P)).v; }

}  // namespace zkb

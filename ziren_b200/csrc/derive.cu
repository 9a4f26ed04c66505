// K7 launchers: see derive.cuh.  One thread per (receive, row) in the build and per (send, row) in the probe; the kernels are
// gather / atomic-scatter passes over resident columns (HBM-bound: every sender column a send touches is read once, coalesced).
#include "derive.h"
#include <stdexcept>

namespace zkb {

constexpr int DV_THREADS = 256;

__global__ void __launch_bounds__(DV_THREADS) derive_fill_kernel(u32* slots, size_t n_slots, u64* counts, size_t n_counts) {
  const size_t i = (size_t)blockIdx.x * DV_THREADS + threadIdx.x;
  if (i < n_slots) slots[i] = DERIVE_EMPTY;
  if (i < n_counts) counts[i] = 0;
}
__global__ void __launch_bounds__(DV_THREADS) derive_build_kernel(u32 n_recv, const DeriveReceive* recv, const DevLookup* lookups,
                                                                  const DevVPC* vpcs, const DevTerm* terms, DeriveTable r, u32* slots,
                                                                  u32 cap_mask) {
  const size_t i = (size_t)blockIdx.x * DV_THREADS + threadIdx.x;
  if (i >= (size_t)n_recv * r.height) return;
  derive_insert((u32)(i / r.height), i % r.height, recv, lookups, vpcs, terms, r, slots, cap_mask);
}
__global__ void __launch_bounds__(DV_THREADS) derive_probe_kernel(u32 send_lookup, DeriveTable s_tab, const DeriveReceive* recv,
                                                                  const DevLookup* lookups, const DevVPC* vpcs, const DevTerm* terms,
                                                                  DeriveTable r, const u32* slots, u32 cap_mask, u64* counts, u64* stats) {
  const size_t row = (size_t)blockIdx.x * DV_THREADS + threadIdx.x;
  if (row >= s_tab.height) return;
  const int rc = derive_probe(lookups[send_lookup], row, s_tab, recv, lookups, vpcs, terms, r, slots, cap_mask, counts);
  if (rc) ZKB_ATOMIC_ADD_U64(stats + (rc - 1), 1);
}
__global__ void __launch_bounds__(DV_THREADS) derive_finish_kernel(const u64* counts, size_t n, u32* out) {
  const size_t i = (size_t)blockIdx.x * DV_THREADS + threadIdx.x;
  if (i < n) out[i] = derive_finish(counts[i]);
}

// which receives of a chip can be derived: every value made of preprocessed columns (and constants), the multiplicity one
// main column with weight one
static bool derivable(const HostLookup& l, u32& mult_col) {
  if (l.is_send || l.scope != 0 || l.values.size() > DERIVE_MAX_VALUES) return false;      // local-scope receives only
  for (auto& v : l.values)
    for (auto& t : v.terms) if (t.is_main) return false;
  if (l.mult.const_canon != 0 || l.mult.terms.size() != 1 || !l.mult.terms[0].is_main || l.mult.terms[0].w_canon != 1) return false;
  mult_col = l.mult.terms[0].col;
  return true;
}

u64 derive_multiplicities(const MachineInfo& m, const ChipInfo& receiver, const u32* receiver_prep, size_t receiver_height,
                          const std::vector<DeriveSender>& senders, u32* out, cudaStream_t s) {
  if (!receiver_height || receiver_height * 64 > 0xffffffffull) throw std::runtime_error("zkb200: derive_multiplicities: bad receiver height");
  std::vector<DeriveReceive> recv;
  std::vector<u32> kinds;
  for (size_t i = 0; i < receiver.lookups.size(); i++) {
    u32 col;
    if (!derivable(receiver.lookups[i], col)) continue;
    recv.push_back({receiver.dev_lookup_begin + (u32)i, col});
    kinds.push_back(receiver.lookups[i].kind);
  }
  if (recv.empty() || recv.size() > 64) throw std::runtime_error("zkb200: derive_multiplicities: " + receiver.name + " has no receive of preprocessed columns with a main column as its multiplicity");
  const size_t entries = recv.size() * receiver_height;
  size_t cap = 1;
  while (cap < 2 * entries) cap <<= 1;
  const size_t n_counts = (size_t)receiver.main_width * receiver_height;
  DevBuf d_recv(2 * recv.size(), s), d_slots(cap, s), d_counts(2 * n_counts, s), d_stats(4, s);
  ZKB_CUDA(cudaMemcpyAsync(d_recv.p, recv.data(), recv.size() * sizeof(DeriveReceive), cudaMemcpyHostToDevice, s));
  ZKB_CUDA(cudaMemsetAsync(d_stats.p, 0, 16, s));
  u64* counts = reinterpret_cast<u64*>(d_counts.p);
  u64* stats = reinterpret_cast<u64*>(d_stats.p);
  const DeriveReceive* recv_dev = reinterpret_cast<const DeriveReceive*>(d_recv.p);
  const size_t fill = cap > n_counts ? cap : n_counts;
  derive_fill_kernel<<<(unsigned)ceil_div(fill, (size_t)DV_THREADS), DV_THREADS, 0, s>>>(d_slots.p, cap, counts, n_counts);
  ZKB_CHECK_LAUNCH();
  const DeriveTable r{receiver_prep, nullptr, receiver_height};
  derive_build_kernel<<<(unsigned)ceil_div(entries, (size_t)DV_THREADS), DV_THREADS, 0, s>>>((u32)recv.size(), recv_dev, m.d_lookups, m.d_vpcs,
                                                                                              m.d_terms, r, d_slots.p, (u32)(cap - 1));
  ZKB_CHECK_LAUNCH();
  for (auto& snd : senders) {
    if (!snd.table.height) continue;
    for (size_t i = 0; i < snd.chip->lookups.size(); i++) {
      const HostLookup& l = snd.chip->lookups[i];
      bool wanted = l.is_send && l.scope == 0 && l.values.size() <= DERIVE_MAX_VALUES;
      if (wanted) { wanted = false; for (u32 k : kinds) wanted = wanted || k == l.kind; }
      if (!wanted) continue;
      derive_probe_kernel<<<(unsigned)ceil_div(snd.table.height, (size_t)DV_THREADS), DV_THREADS, 0, s>>>(
          snd.chip->dev_lookup_begin + (u32)i, snd.table, recv_dev, m.d_lookups, m.d_vpcs, m.d_terms, r, d_slots.p, (u32)(cap - 1), counts, stats);
      ZKB_CHECK_LAUNCH();
    }
  }
  derive_finish_kernel<<<(unsigned)ceil_div(n_counts, (size_t)DV_THREADS), DV_THREADS, 0, s>>>(counts, n_counts, out);
  ZKB_CHECK_LAUNCH();
  u64 h_stats[2] = {0, 0};
  ZKB_CUDA(cudaMemcpyAsync(h_stats, stats, 16, cudaMemcpyDeviceToHost, s));
  ZKB_CUDA(cudaStreamSynchronize(s));
  if (h_stats[1]) throw std::runtime_error("zkb200: derive_multiplicities: " + std::to_string(h_stats[1]) + " lookups are in no row of " + receiver.name);
  return h_stats[0];
}

}  // namespace zkb

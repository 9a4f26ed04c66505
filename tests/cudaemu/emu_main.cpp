// Entry points of the emulated build (tests/cudaemu/cuda_runtime.h): the product's own launchers, compiled from the rewritten
// csrc/tracegen.cu, csrc/derive.cu and csrc/machine.cpp, behind a C ABI for tests/test_cuda_emulation.py.
#include "derive.h"
#include "tracegen.h"
#include <cstring>
#include <string>

namespace zkb { std::atomic<unsigned long long> g_kernel_launches{0}; }
using namespace zkb;

static thread_local std::string emu_error;
template <class F> static int guarded(F f) {
  try { f(); return 0; } catch (const std::exception& e) { emu_error = e.what(); return 1; }
}
extern "C" {
const char* emu_last_error() { return emu_error.c_str(); }
unsigned long long emu_launches() { return g_kernel_launches.load(); }
int emu_chip_width(const char* chip) {
  if (!strcmp(chip, "KeccakSponge")) return KS_WIDTH;
  if (!strcmp(chip, "Global")) return GLOBAL_WIDTH;
  const int id = alu_chip_by_name(chip);
  return id < 0 ? -1 : alu_width(id);
}
// what zkb200_generate_alu_trace / zkb200_generate_keccak_sponge_trace call (csrc/capi.cu) after staging the events
int emu_generate_trace(const char* chip, const uint32_t* events, size_t n_events, size_t height, uint32_t* out, int col_major) {
  return guarded([&] {
    static const bool once = (tracegen_upload_constants(), true);
    (void)once;
    if (!strcmp(chip, "KeccakSponge")) {
      if (!col_major) throw std::runtime_error("the KeccakSponge kernel writes column-major only");
      keccak_sponge_trace(events, n_events, height, out, nullptr);
    } else if (!strcmp(chip, "Global")) global_trace(events, n_events, height, out, col_major != 0, nullptr);
    else alu_trace(alu_chip_by_name(chip), events, n_events, height, out, col_major != 0, nullptr);
  });
}
// what zkb200_derive_multiplicities calls (csrc/capi.cu): machine tables from the descriptor as zkb200_ctx_create builds them
int emu_derive(const uint32_t* desc, size_t n_words, const char* receiver, const uint32_t* receiver_prep, size_t receiver_height,
               int n_senders, const char* const* names, const uint32_t* const* preps, const uint32_t* const* mains,
               const size_t* heights, uint32_t* out, unsigned long long* n_lookups) {
  return guarded([&] {
    MachineInfo m;
    m.parse(desc, n_words);
    m.upload();
    struct Cleanup { MachineInfo& m; ~Cleanup() { m.destroy(); } } cleanup{m};
    const ChipInfo* r = m.find(receiver);
    if (!r) throw std::runtime_error("unknown chip");
    std::vector<DeriveSender> snd;
    for (int i = 0; i < n_senders; i++) {
      const ChipInfo* c = m.find(names[i]);
      if (!c) throw std::runtime_error("unknown chip");
      snd.push_back({c, DeriveTable{preps[i], mains[i], heights[i]}});
    }
    *n_lookups = derive_multiplicities(m, *r, receiver_prep, receiver_height, snd, out, nullptr);
  });
}
}

// Compiled host side above the C ABI: proves the shards of a case file through include/zkb200.hpp
// (the C++ mirror of the MachineProver trait) and writes the proofs, so that the drop-in boundary is
// exercised from compiled code without Python in the process.
//   prove_shard <case.zkcase> <proofs.out> [device]
// Case file (little-endian u32 words unless noted; written by ziren_b200/casefile.py):
//   "ZKCS", 1, n_desc, desc[n_desc], pc_start, init_global_sum[14],
//   n_prep, n_prep x trace, n_records, per record: n_traces, n_traces x trace, n_pv, pv[n_pv]
//   trace = name_len, name bytes padded to words, height (u64 as two words), width, data[height*width] (Montgomery)
// Output: "ZKPO", 1, preprocessed_commit[8], n_records, per record: n_words, proof words (ZKPF).
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include "../include/zkb200.hpp"

namespace {
struct Reader {
  std::vector<uint32_t> w;
  size_t at = 0;
  uint32_t u32() { if (at >= w.size()) throw std::runtime_error("case file truncated"); return w[at++]; }
  const uint32_t* take(size_t n) { if (at + n > w.size()) throw std::runtime_error("case file truncated"); const uint32_t* p = w.data() + at; at += n; return p; }
  zkb200::Trace trace() {
    const uint32_t len = u32();
    const uint32_t* nb = take((len + 3) / 4);
    zkb200::Trace t;
    t.name.assign(reinterpret_cast<const char*>(nb), len);
    const uint64_t lo = u32(), hi = u32();
    t.height = (size_t)(lo | (hi << 32));
    t.width = u32();
    t.data = take(t.height * t.width);
    return t;
  }
};
}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) { std::fprintf(stderr, "usage: %s <case.zkcase> <proofs.out> [device]\n", argv[0]); return 2; }
  try {
    std::ifstream f(argv[1], std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error(std::string("cannot open ") + argv[1]);
    const std::streamsize bytes = f.tellg();
    f.seekg(0);
    Reader r;
    r.w.resize((size_t)bytes / 4);
    f.read(reinterpret_cast<char*>(r.w.data()), (std::streamsize)(r.w.size() * 4));
    if (r.u32() != 0x53434b5au || r.u32() != 1) throw std::runtime_error("not a ZKCS v1 case file");
    const uint32_t n_desc = r.u32();
    const uint32_t* d = r.take(n_desc);
    std::vector<uint32_t> desc(d, d + n_desc);
    const uint32_t pc_start = r.u32();
    std::array<uint32_t, 14> gsum{};
    std::memcpy(gsum.data(), r.take(14), 14 * 4);
    std::vector<zkb200::Trace> prep;
    for (uint32_t i = 0, n = r.u32(); i < n; i++) prep.push_back(r.trace());
    std::vector<std::pair<std::vector<zkb200::Trace>, std::vector<uint32_t>>> records;
    for (uint32_t i = 0, n = r.u32(); i < n; i++) {
      std::vector<zkb200::Trace> tr;
      for (uint32_t j = 0, m = r.u32(); j < m; j++) tr.push_back(r.trace());
      const uint32_t npv = r.u32();
      const uint32_t* pv = r.take(npv);
      records.emplace_back(std::move(tr), std::vector<uint32_t>(pv, pv + npv));
    }

    zkb200::B200Prover prover(argc > 3 ? std::atoi(argv[3]) : 0, desc);     // MachineProver::new
    zkb200::ProvingKey pk = prover.setup(prep, pc_start, &gsum);            // setup + pk_to_device
    auto proofs = prover.prove(pk, records);                                // commit + open per record

    std::ofstream o(argv[2], std::ios::binary);
    auto put = [&](uint32_t v) { o.write(reinterpret_cast<const char*>(&v), 4); };
    put(0x4f504b5au); put(1);
    for (uint32_t v : pk.preprocessed_commit()) put(v);
    put((uint32_t)proofs.size());
    for (auto& p : proofs) { put((uint32_t)p.size()); o.write(reinterpret_cast<const char*>(p.data()), (std::streamsize)(p.size() * 4)); }
    std::printf("proved %zu shard(s)\n", proofs.size());
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "prove_shard: %s\n", e.what());
    return 1;
  }
}

// zkb200.hpp — C++17 host-side mirror of the reference's `MachineProver` plugin trait
// (crates/stark/src/prover.rs:30-184) over the C ABI of zkb200.h.  The reference's host code is
// compiled Rust (no Rust toolchain exists in this build image), so this is the
// compiled-language host side above the ABI: same method names, argument meaning and error
// behaviour as the trait, RAII for the device objects the trait moves around by value.  It holds
// no logic of its own beyond marshalling; INTEGRATION.md shows the equivalent Rust shim.
//
//   trait item (prover.rs)                          here
//   MachineProver::new(machine)            :43      B200Prover(device, descriptor)
//   setup(&Program) / pk_to_device         :49-63   setup(preprocessed traces, pc_start, initial_global_cumulative_sum)
//   generate_traces(record)                :70-108  generate_trace(chip, events, log_height)   (ALU / control-flow chips),
//                                                   generate_keccak_sponge_trace(blocks, ...), events_trace(...) into commit()
//   commit(record, traces)                 :258     commit(traces, public_values) -> ShardMainData
//   open(pk, data, &mut challenger)        :298     open(pk, data, challenger) -> proof words (ZKPF)
//   prove(pk, records, challenger, opts)   :660-693 prove(pk, records)
//   MachineProvingKey::observe_into        :714     ProvingKey::observe_into()
//   Self::Error                            :206     zkb200::Error (std::runtime_error with zkb200_last_error)
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
#include "zkb200.h"

namespace zkb200 {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// RowMajorMatrix<KoalaBear> as the trait's `traces: Vec<(String, RowMajorMatrix<Val>)>` carries it:
// Montgomery words, row-major; `data` may be host or device memory and is borrowed.
struct Trace {
  std::string name;
  const uint32_t* data = nullptr;
  size_t height = 0, width = 0;
  uint32_t flags = 0;        // 0, ZKB200_TRACE_COL_MAJOR or ZKB200_TRACE_EVENTS (zkb200.h)
  size_t n_events = 0;       // ZKB200_TRACE_EVENTS: event records behind `data`
};
using Commitment = std::array<uint32_t, 8>;    // canonical
using Challenger = std::array<uint32_t, 34>;   // DuplexChallenger image, canonical (zkb200.h)

class B200Prover;

// MachineProver::DeviceProvingKey (preprocessed traces + LDEs + tree on the GPU)
class ProvingKey {
 public:
  ProvingKey() = default;
  ProvingKey(ProvingKey&& o) noexcept : h_(o.h_), commit_(o.commit_) { o.h_ = nullptr; }
  ProvingKey& operator=(ProvingKey&& o) noexcept { if (this != &o) { reset(); h_ = o.h_; commit_ = o.commit_; o.h_ = nullptr; } return *this; }
  ProvingKey(const ProvingKey&) = delete;
  ProvingKey& operator=(const ProvingKey&) = delete;
  ~ProvingKey() { reset(); }
  const Commitment& preprocessed_commit() const { return commit_; }
  // the challenger every shard proof starts from: a fresh one after pk.observe_into (prover.rs:714-721)
  Challenger observe_into() const {
    Challenger c{};
    if (zkb200_pk_initial_challenger(h_, c.data())) throw Error(zkb200_last_error(nullptr));
    return c;
  }
  const zkb200_pk* handle() const { return h_; }

 private:
  friend class B200Prover;
  void reset() { if (h_) zkb200_pk_free(h_); h_ = nullptr; }
  zkb200_pk* h_ = nullptr;
  Commitment commit_{};
};

// ShardMainData<SC, DeviceMatrix, DeviceProverData> (crates/stark/src/types.rs:16-22): moved into open()
class ShardMainData {
 public:
  ShardMainData() = default;
  ShardMainData(ShardMainData&& o) noexcept : main_commit(o.main_commit), h_(o.h_) { o.h_ = nullptr; }
  ShardMainData& operator=(ShardMainData&& o) noexcept { if (this != &o) { reset(); h_ = o.h_; main_commit = o.main_commit; o.h_ = nullptr; } return *this; }
  ShardMainData(const ShardMainData&) = delete;
  ShardMainData& operator=(const ShardMainData&) = delete;
  ~ShardMainData() { reset(); }
  Commitment main_commit{};

 private:
  friend class B200Prover;
  void reset() { if (h_) zkb200_shard_free(h_); h_ = nullptr; }
  zkb200_shard* h_ = nullptr;
};

// device matrix produced by trace generation; usable as a Trace for commit()
class DeviceTrace {
 public:
  DeviceTrace(std::string name, uint32_t* dev, size_t height, size_t width, void (*free_fn)(void*))
      : name_(std::move(name)), dev_(dev), h_(height), w_(width), free_(free_fn) {}
  DeviceTrace(DeviceTrace&& o) noexcept : name_(std::move(o.name_)), dev_(o.dev_), h_(o.h_), w_(o.w_), free_(o.free_) { o.dev_ = nullptr; }
  DeviceTrace(const DeviceTrace&) = delete;
  ~DeviceTrace() { if (dev_ && free_) free_(dev_); }
  Trace view() const { return Trace{name_, dev_, h_, w_}; }

 private:
  std::string name_;
  uint32_t* dev_;
  size_t h_, w_;
  void (*free_)(void*);
};

// `impl MachineProver<KoalaBearPoseidon2, A> for B200Prover`: one instance per GPU; commit/open may be
// called from several host threads (crates/core/machine/src/utils/prove.rs:487-521 does).
class B200Prover {
 public:
  // `machine` is the ZKMD descriptor of the StarkMachine (zkb200.h)
  B200Prover(int device, const std::vector<uint32_t>& machine) {
    if (zkb200_ctx_create(device, machine.data(), machine.size(), &ctx_)) throw Error(zkb200_last_error(nullptr));
  }
  B200Prover(const B200Prover&) = delete;
  B200Prover& operator=(const B200Prover&) = delete;
  ~B200Prover() { if (ctx_) zkb200_ctx_destroy(ctx_); }

  ProvingKey setup(const std::vector<Trace>& preprocessed, uint32_t pc_start = 0,
                   const std::array<uint32_t, 14>* initial_global_cumulative_sum = nullptr) const {
    // nullptr: SepticDigest::zero(), the curve START point (crates/stark/src/septic_digest.rs:9-42)
    auto t = marshal(preprocessed);
    ProvingKey pk;
    check(zkb200_setup(ctx_, t.data(), (int)t.size(), pc_start,
                       initial_global_cumulative_sum ? initial_global_cumulative_sum->data() : nullptr, pk.commit_.data(), &pk.h_));
    return pk;
  }
  ShardMainData commit(const std::vector<Trace>& traces, const std::vector<uint32_t>& public_values) const {
    auto t = marshal(traces);
    ShardMainData d;
    check(zkb200_commit(ctx_, t.data(), (int)t.size(), public_values.data(), public_values.size(), d.main_commit.data(), &d.h_));
    return d;
  }
  // consumes `data`; `challenger` is the per-shard clone and is advanced like the trait's `&mut`
  std::vector<uint32_t> open(const ProvingKey& pk, ShardMainData data, Challenger& challenger) const {
    uint32_t* words = nullptr;
    size_t n = 0;
    check(zkb200_open(ctx_, pk.handle(), data.h_, challenger.data(), &words, &n));
    std::vector<uint32_t> proof(words, words + n);
    zkb200_free(words);
    return proof;
  }
  // prove(pk, records, ..): every shard starts from a clone of the post-observe_into challenger
  std::vector<std::vector<uint32_t>> prove(const ProvingKey& pk,
                                           const std::vector<std::pair<std::vector<Trace>, std::vector<uint32_t>>>& records) const {
    const Challenger base = pk.observe_into();
    std::vector<std::vector<uint32_t>> proofs;
    for (auto& rec : records) {
      Challenger ch = base;
      proofs.push_back(open(pk, commit(rec.first, rec.second), ch));
    }
    return proofs;
  }
  // generate_trace of an ALU / control-flow chip from the record's event vector (28-byte records, host or
  // device); `out_dev` is caller-provided device memory of 2^log_height x width words
  void generate_trace(const std::string& chip, const void* events, size_t n_events, unsigned log_height, uint32_t* out_dev,
                      bool col_major = false) const {
    check(zkb200_generate_alu_trace(ctx_, chip.c_str(), events, n_events, log_height, out_dev, col_major ? 1 : 0));
  }
  // generate_trace of the KeccakSponge chip from flattened block records (zkb200_keccak_block, 24 rows each)
  void generate_keccak_sponge_trace(const zkb200_keccak_block* blocks, size_t n_blocks, unsigned log_height, uint32_t* out_dev,
                                    bool col_major = false) const {
    check(zkb200_generate_keccak_sponge_trace(ctx_, blocks, n_blocks, log_height, out_dev, col_major ? 1 : 0));
  }
  // a table handed to commit() as the chip's event records: the row filler runs inside the commit (ZKB200_TRACE_EVENTS)
  static Trace events_trace(const std::string& chip, const void* events, size_t n_events, unsigned log_height, size_t width) {
    Trace t;
    t.name = chip; t.data = static_cast<const uint32_t*>(events); t.height = (size_t)1 << log_height; t.width = width;
    t.flags = ZKB200_TRACE_EVENTS; t.n_events = n_events;
    return t;
  }
  static int trace_width(const std::string& chip) {
    return chip == "KeccakSponge" ? zkb200_keccak_sponge_trace_width() : zkb200_alu_trace_width(chip.c_str());
  }
  void sync() const { check(zkb200_sync(ctx_)); }
  zkb200_ctx* handle() const { return ctx_; }

 private:
  void check(int rc) const { if (rc) throw Error(zkb200_last_error(ctx_)); }
  static std::vector<zkb200_trace> marshal(const std::vector<Trace>& v) {
    std::vector<zkb200_trace> t;
    t.reserve(v.size());
    for (auto& x : v) t.push_back(zkb200_trace{x.name.c_str(), x.data, x.height, x.width, x.flags, x.n_events});
    return t;
  }
  zkb200_ctx* ctx_ = nullptr;
};

}  // namespace zkb200

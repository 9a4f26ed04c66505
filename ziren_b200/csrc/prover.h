// Host orchestration of the shard prover on one GPU: the device-side counterpart of the
// reference's MachineProver implementation (crates/stark/src/prover.rs:30-184 trait,
// :258-292 commit, :298-653 open) and of StarkMachine::setup (crates/stark/src/machine.rs:352-459).
#pragma once
#include "lane_pool.h"
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "common.h"
#include "fri.h"
#include "hash.h"
#include "machine.h"
#include "ntt.h"

namespace zkb {

// One committed batch of matrices: the device image of Plonky3's PcsProverData
// (bit-reversed coset LDEs + Merkle digest layers).
struct Commit {
  std::vector<DevMat> ldes;          // column-major, height n << log_blowup
  std::vector<unsigned> log_n;       // trace log-heights
  DigestLayers layers;
  u32 root[8] = {0};                 // Montgomery
  unsigned log_max_height = 0;       // of the LDEs
  bool empty() const { return ldes.empty(); }
};

// A compute lane: one stream with its own scratch.  commit/open calls from different host threads
// run on different lanes, so the phases of one shard that leave the SMs idle (FRI commit phase,
// grinding, host round trips) and kernels bound by different pipes overlap with another shard's.
struct Lane {
  cudaStream_t stream = nullptr;
  ParamArena arena;
  u32* d_small = nullptr;            // small device scratch (roots, sums)
  u32* h_small = nullptr;            // pinned mirror
  u32* h_big = nullptr;              // pinned, device-mapped staging for host -> device tables (proof skeleton, gather jobs)
  size_t h_big_bytes = (size_t)32 << 20;
  std::mutex mu;
  KeepAlive keep;                    // shared tables the queued kernels read (ntt.h)
  // start of a call that owns the lane: the previous call drained the stream before it returned
  void begin() { arena.reset(); keep.clear(); }
};
constexpr int NUM_LANES = 8;
enum UploadMode { UPLOAD_PULL = 0, UPLOAD_DMA = 1, UPLOAD_DMA2D = 2 };

// stream-ordered wait on a 32-bit device counter (>=): the driver's cuStreamWaitValue32, fetched at run
// time so that the library links against the runtime only.  Unavailable -> operator bool is false.
struct StreamWaitValue {
  typedef int (*Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
  Fn fn = nullptr;
  void init();
  bool probe(cudaStream_t s, const u32* zeroed_counter) const;
  explicit operator bool() const { return fn != nullptr; }
  void operator()(cudaStream_t s, const u32* addr, u32 value) const {
    const int rc = fn(s, (unsigned long long)(uintptr_t)addr, value, 0x0 /* CU_STREAM_WAIT_VALUE_GEQ */);
    if (rc != 0) throw std::runtime_error("zkb200: cuStreamWaitValue32 failed with CUresult " + std::to_string(rc));
  }
};

// Pageable host memory -> device through a ring of pinned slots filled by a few host threads
// (prover.cu).  Lazily allocated: a caller that hands over pinned or device memory never pays for it.
struct HostStager {
  int nslots = 4, nthreads = 8;
  size_t slot_bytes = (size_t)32 << 20;
  std::vector<u32*> slots;
  std::vector<cudaEvent_t> free_ev;
  int next_slot = 0;
  std::mutex use_mu;
  struct Job { u32* dst; const u32* src; size_t row_words, src_pitch, rows; };
  std::vector<std::thread> workers;
  std::mutex m;
  std::condition_variable cv, cv_done;
  Job job{};
  unsigned long long job_id = 0;
  int done = 0;
  bool stop = false;
  void init(int nslots, size_t slot_bytes, int nthreads);
  void ensure();
  void destroy();
  void worker(int t);
  void copy_2d(u32* dst, const u32* src, size_t row_words, size_t src_pitch, size_t rows, cudaStream_t s);
};

struct Ctx {
  int device = 0;
  Lane lanes[NUM_LANES];
  // More lanes than host threads in flight: a commit holds its lane while its upload is still queued
  // behind another shard's, and an open() must not have to wait for such an idle lane (3 lanes and 4
  // threads starved the opens: 151 ms per shard).  ZKB200_LANES=1 serialises all compute on one stream.
  int active_lanes = 8;
  cudaStream_t copy_stream = nullptr;   // host->device uploads + layout change, overlaps the compute lanes
  std::mutex copy_mu;
  HostStager stager;                    // pageable host sources
  size_t piece_bytes = (size_t)256 << 20;   // column pieces of the main commit (prover_commit)
  // pinned host traces (ZKB200_UPLOAD=pull|dma|dma2d), see prover_commit.  Measured on the bench shard (ms per
  // shard, 4 shards / 1 shard in flight): pull 121.5 / 149.3, dma 128.1 / 193.1, dma2d 247.2 / 278.7.  The pull
  // kernel moves 48.6 GB/s alone and about 40 GB/s next to the compute kernels, but it lets the LDE and the leaf
  // hashing of a shard run under that shard's own upload; the contiguous DMA (55 GB/s) cannot.
  int upload_mode = UPLOAD_PULL;
  int pull_ctas = 32;                   // persistent CTAs of the pull kernel (ZKB200_PULL_CTAS)
  bool pull_split = true;               // pull mode: single-piece matrices go by contiguous DMA on `dma_stream` (ZKB200_PULL_SPLIT=0: all pulled)
  cudaStream_t dma_stream = nullptr;
  bool pull_exclusive = false;          // 1024-thread CTAs that own their SM (ZKB200_PULL_EXCLUSIVE=1, 8 CTAs by default)
  StreamWaitValue wait_value;           // cuStreamWaitValue32: a lane waits for a counter of the pull kernel
  static constexpr size_t PULL_COUNTER_RING = 1 << 16;
  u32* pull_counters = nullptr;         // ring of counters (cudaMalloc), slices handed out under copy_mu
  size_t pull_counter_next = 0;
  LanePool<NUM_LANES> pool;             // which lane a commit/open call runs on
  MachineInfo machine;
  NttTables tables;
  // statistics of the last open(): kernel-stage timings (ms) when profiling is enabled
  bool profile = false;
  std::vector<std::pair<std::string, float>> stage_ms;
  std::mutex stage_mu;                  // stage_ms is appended to from every lane
  unsigned long long launches = 0;

  void init(int device, const u32* desc, size_t n);
  void destroy();
};

// row-major Montgomery, host or device pointer (flags 0); device column-major (flags 1); event records (flags 2:
// include/zkb200.h ZKB200_TRACE_COL_MAJOR / ZKB200_TRACE_EVENTS)
struct TraceIn { std::string name; const u32* data; size_t height, width; u32 flags = 0; size_t n_events = 0; };
constexpr u32 TRACE_COL_MAJOR = 1u, TRACE_EVENTS = 2u, TRACE_DERIVED = 4u;

struct Pk {
  Ctx* ctx = nullptr;
  std::vector<std::string> names;      // (height desc, name asc)
  std::vector<DevMat> traces;          // preprocessed traces, column-major
  std::vector<bool> local_only;
  Commit data;
  u32 commit_canon[8] = {0};
  u32 pc_start = 0;                    // canonical
  u32 init_global_sum[14] = {0};       // canonical
  int index_of(const std::string& n) const {
    for (size_t i = 0; i < names.size(); i++) if (names[i] == n) return (int)i;
    return -1;
  }
};

struct Shard {
  Ctx* ctx = nullptr;
  std::vector<std::string> names;      // (height desc, name asc)
  std::vector<DevMat> traces;          // main traces, column-major
  Commit main;
  std::vector<u32> public_values;      // canonical
};

Pk* prover_setup(Ctx& ctx, const std::vector<TraceIn>& prep, u32 pc_start, const u32* init_gsum);
Shard* prover_commit(Ctx& ctx, const std::vector<TraceIn>& traces, const u32* pv, size_t npv);
// commit of a shard with ZKB200_TRACE_DERIVED tables (multiplicity columns derived from the shard's other tables, K7): needs the
// proving key for the preprocessed tables, so only zkb200_prove_shard reaches it
Shard* prover_commit_derived(Ctx& ctx, const Pk& pk, const std::vector<TraceIn>& traces, const u32* pv, size_t npv);
// consumes nothing; caller frees the shard.  challenger34: canonical image, in/out.
std::vector<u32> prover_open(Ctx& ctx, const Pk& pk, Shard& shard, u32* challenger34);

// building blocks shared with the micro entry points
DevMat upload_colmajor(Ctx& ctx, const u32* data, size_t h, size_t w, cudaStream_t on, cudaStream_t free_on);
void pcs_commit(Ctx& ctx, Lane& L, std::vector<DevMat>& traces, const std::vector<Fp>& domain_shifts, Commit& out,
                const char* lde_stage, const char* merkle_stage);

// takes a free lane, waiting for any of them when all are busy, and holds it.  The lane's own mutex
// is held as well: the kernel-level entry points (capi.cu) lock lanes[0].mu directly.
struct LaneGuard {
  Ctx& ctx;
  int index;
  Lane* lane;
  explicit LaneGuard(Ctx& c) : ctx(c), index(c.pool.acquire(c.active_lanes)), lane(&c.lanes[index]) { lane->mu.lock(); }
  ~LaneGuard() {
    lane->mu.unlock();
    ctx.pool.release(index);
  }
  LaneGuard(const LaneGuard&) = delete;
  LaneGuard& operator=(const LaneGuard&) = delete;
};

}  // namespace zkb

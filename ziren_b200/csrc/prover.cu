// Shard prover orchestration on one GPU (see prover.h).  All numeric work is in the kernels of
// ntt.cu / hash.cu / logup.cu / quotient.cu / open.cu / fri.cu; this file sequences them on the
// context stream, runs the duplex challenger on the host between stages (the transcript order of
// crates/stark/src/prover.rs:298-653) and packs the "ZKPF" proof.
#include "prover.h"
#include "derive.h"
#include "tracegen.h"
#include <algorithm>
#include <array>
#include <cstdlib>
#include <map>
#include <nvtx3/nvToolsExt.h>
#include "layout.h"
#include "logup.h"
#include "open.h"
#include "quotient.h"

namespace zkb {

std::atomic<unsigned long long> g_kernel_launches{0};

// CURVE_CUMULATIVE_SUM_START_{X,Y} (canonical), crates/stark/src/septic_digest.rs:9-14
static const u32 SEPTIC_DIGEST_ZERO[14] = {637514027u, 1595065213u, 1998064738u, 72333738u, 1211544370u, 822986770u, 1518535784u,
                                           1604177449u, 90440090u, 259343427u, 140470264u, 1162099742u, 941559812u, 1064053343u};

void Ctx::init(int dev, const u32* desc, size_t n) {
  device = dev;
  if (const char* e = getenv("ZKB200_LANES")) active_lanes = std::max(1, std::min(NUM_LANES, atoi(e)));
  ZKB_CUDA(cudaSetDevice(dev));
  machine.parse(desc, n);
  for (auto& L : lanes) ZKB_CUDA(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
  {
    // the copy stream's "pull" kernels must get their few CTAs even while the lanes fill the SMs
    int lo = 0, hi = 0;
    ZKB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    ZKB_CUDA(cudaStreamCreateWithPriority(&copy_stream, cudaStreamNonBlocking, hi));
  }
  if (const char* e = getenv("ZKB200_UPLOAD")) upload_mode = std::string(e) == "dma" ? UPLOAD_DMA : std::string(e) == "dma2d" ? UPLOAD_DMA2D : UPLOAD_PULL;
  // counters the pull kernel bumps and the lanes wait on: plain cudaMalloc memory (stream memory
  // operations do not take stream-ordered pool allocations), handed out as a ring; probed once here -
  // a driver that refuses the wait turns the pull mode off (2-D DMA instead)
  wait_value.init();
  ZKB_CUDA(cudaMalloc((void**)&pull_counters, PULL_COUNTER_RING * sizeof(u32)));
  ZKB_CUDA(cudaMemset(pull_counters, 0, PULL_COUNTER_RING * sizeof(u32)));
  if (wait_value && !wait_value.probe(lanes[0].stream, pull_counters)) wait_value.fn = nullptr;
  if (const char* e = getenv("ZKB200_PULL_SPLIT")) pull_split = atoi(e) != 0;
  {
    int lo = 0, hi = 0;
    ZKB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    ZKB_CUDA(cudaStreamCreateWithPriority(&dma_stream, cudaStreamNonBlocking, hi));
  }
  if (const char* e = getenv("ZKB200_PULL_EXCLUSIVE")) pull_exclusive = atoi(e) != 0;
  if (pull_exclusive) pull_ctas = 8;
  if (const char* e = getenv("ZKB200_PULL_CTAS")) pull_ctas = std::max(1, atoi(e));
  pull_set_device_attributes();
  // keep freed blocks in the stream-ordered pool: shard proofs reuse the same sizes
  cudaMemPool_t pool;
  ZKB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
  unsigned long long thr = ~0ull;
  ZKB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
  // never let the pool make the copy stream wait for the compute stream (or vice versa) just to
  // recycle a block: a fresh block is cheaper than losing the upload/compute overlap
  int no = 0;
  ZKB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &no));
  // Prime the pool: growing it (cuMemCreate/cuMemMap of GBs) in the middle of a proof costs up to
  // seconds, so reserve the physical memory once.  ZKB200_POOL_GB overrides (0 disables).
  {
    size_t free_b = 0, total_b = 0;
    ZKB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    // default: 96 GB or 55 % of what is free - three or four 2^20..2^21-row shards in flight (about 25 GB each);
    // a smaller reservation (32 GB was tried) makes the pool grow in the middle of timed proofs (value arm
    // 169 ms instead of 109 ms per shard).  ZKB200_POOL_GB sets it for a GPU shared with other tenants.
    size_t want = std::min<size_t>((size_t)96 << 30, free_b / 20 * 11);
    if (const char* e = getenv("ZKB200_POOL_GB")) want = std::min<size_t>((size_t)atol(e) << 30, free_b * 9 / 10);
    if (want) {
      void* p = nullptr;
      ZKB_CUDA(cudaMallocAsync(&p, want, lanes[0].stream));
      ZKB_CUDA(cudaFreeAsync(p, lanes[0].stream));
      ZKB_CUDA(cudaStreamSynchronize(lanes[0].stream));
    }
  }
  {
    int nthr = getenv("ZKB200_STAGE_THREADS") ? atoi(getenv("ZKB200_STAGE_THREADS")) : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
    int mb = getenv("ZKB200_STAGE_SLOT_MB") ? atoi(getenv("ZKB200_STAGE_SLOT_MB")) : 32;
    stager.init(4, (size_t)std::max(1, mb) << 20, std::max(1, nthr));
  }
  if (const char* e = getenv("ZKB200_PIECE_MB")) piece_bytes = (size_t)std::max(1, atoi(e)) << 20;
  if (const char* e = getenv("ZKB200_PIECE_KB")) piece_bytes = (size_t)std::max(1, atoi(e)) << 10;   // tests: many pieces at small sizes
  machine.upload();
  tables.init(lanes[0].stream);
  p2_upload_constants();
  tracegen_upload_constants();
  for (auto& L : lanes) {
    L.arena.init(8u << 20, L.stream);
    ZKB_CUDA(cudaMalloc((void**)&L.d_small, 1 << 16));
    ZKB_CUDA(cudaMallocHost((void**)&L.h_small, 1 << 16));
    ZKB_CUDA(cudaHostAlloc((void**)&L.h_big, L.h_big_bytes, cudaHostAllocMapped));
  }
  ZKB_CUDA(cudaStreamSynchronize(lanes[0].stream));
}
void Ctx::destroy() {
  cudaSetDevice(device);
  cudaDeviceSynchronize();
  stager.destroy();
  machine.destroy();
  tables.destroy();
  for (auto& L : lanes) {
    L.arena.destroy();
    if (L.d_small) cudaFree(L.d_small);
    if (L.h_small) cudaFreeHost(L.h_small);
    if (L.h_big) cudaFreeHost(L.h_big);
    L.h_big = nullptr;
    if (L.stream) cudaStreamDestroy(L.stream);
    L.stream = nullptr; L.d_small = nullptr; L.h_small = nullptr;
  }
  if (pull_counters) cudaFree(pull_counters);
  pull_counters = nullptr;
  if (dma_stream) cudaStreamDestroy(dma_stream);
  dma_stream = nullptr;
  if (copy_stream) cudaStreamDestroy(copy_stream);
  copy_stream = nullptr;
}

// Per-stage device times (profiling mode) and NVTX ranges named after the reference's tracing spans
// (crates/stark/src/prover.rs:340,402,427,496,546).  Times of stages with the same name accumulate:
// the main commit interleaves layout change, LDE and leaf hashing piece by piece.
struct StageTimer {
  Ctx& ctx; cudaStream_t stream; const char* name; cudaEvent_t a = nullptr, b = nullptr;
  StageTimer(Ctx& c, Lane& L, const char* n, const char* span = nullptr) : ctx(c), stream(L.stream), name(n) {
    nvtxRangePushA(span ? span : n);
    if (ctx.profile) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, stream); }
  }
  ~StageTimer() {
    if (ctx.profile) {
      cudaEventRecord(b, stream); cudaEventSynchronize(b);
      float ms = 0; cudaEventElapsedTime(&ms, a, b);
      {
        std::lock_guard<std::mutex> lock(ctx.stage_mu);
        bool found = false;
        for (auto& kv : ctx.stage_ms) if (kv.first == name) { kv.second += ms; found = true; break; }
        if (!found) ctx.stage_ms.push_back({name, ms});
      }
      cudaEventDestroy(a); cudaEventDestroy(b);
    }
    nvtxRangePop();
  }
};

bool StreamWaitValue::probe(cudaStream_t s, const u32* zeroed_counter) const {
  if (!fn) return false;
  if (fn(s, (unsigned long long)(uintptr_t)zeroed_counter, 0u, 0x0) != 0) return false;     // 0 >= 0: satisfied at once
  return cudaStreamSynchronize(s) == cudaSuccess;
}
void StreamWaitValue::init() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    fn = reinterpret_cast<Fn>(p);
  else
    cudaGetLastError();
}

static bool is_device_pointer(const void* p, bool* pinned = nullptr) {
  cudaPointerAttributes attr;
  if (pinned) *pinned = false;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return false; }
  if (pinned) *pinned = attr.type == cudaMemoryTypeHost;
  return attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
}

// Row-major host/device matrix -> column-major device matrix, issued on stream `on`.  The result
// is released on the compute stream, so `on` must be joined into it before first use.
DevMat upload_colmajor(Ctx& ctx, const u32* data, size_t h, size_t w, cudaStream_t on, cudaStream_t free_on) {
  DevMat m(h, w, on);
  m.buf.stream = free_on;
  if (h * w == 0) return m;
  bool pinned = false;
  if (is_device_pointer(data, &pinned)) {
    transpose_to_colmajor(data, m.d(), h, w, on);
  } else {
    DevBuf stage(h * w, on);
    if (pinned) ZKB_CUDA(cudaMemcpyAsync(stage.p, data, h * w * sizeof(u32), cudaMemcpyHostToDevice, on));
    else ctx.stager.copy_2d(stage.p, data, w, w, h, on);
    transpose_to_colmajor(stage.p, m.d(), h, w, on);
  }
  return m;
}

// ---- pageable host memory --------------------------------------------------------------------
// The reference hands over RowMajorMatrix.values, an ordinary heap Vec (prover.rs:258-262).  A DMA
// from pageable memory is staged by the driver through one small bounce buffer on the calling
// thread; here a few host threads gather the (strided) rows of a piece into a ring of pinned
// slots while earlier slots are in flight, so the copy runs near the pinned rate.
void HostStager::init(int nslots_, size_t slot_bytes_, int nthreads_) {
  nslots = nslots_; slot_bytes = slot_bytes_; nthreads = nthreads_;
}
void HostStager::ensure() {
  if (!slots.empty()) return;
  slots.resize(nslots); free_ev.resize(nslots);
  for (int i = 0; i < nslots; i++) {
    ZKB_CUDA(cudaMallocHost((void**)&slots[i], slot_bytes));
    ZKB_CUDA(cudaEventCreateWithFlags(&free_ev[i], cudaEventDisableTiming));
  }
  stop = false;
  for (int t = 0; t < nthreads; t++) workers.emplace_back([this, t] { worker(t); });
}
void HostStager::destroy() {
  {
    std::lock_guard<std::mutex> lk(m);
    stop = true;
  }
  cv.notify_all();
  for (auto& w : workers) w.join();
  workers.clear();
  for (auto p : slots) cudaFreeHost(p);
  for (auto e : free_ev) cudaEventDestroy(e);
  slots.clear(); free_ev.clear();
}
void HostStager::worker(int t) {
  unsigned long long seen = 0;
  for (;;) {
    std::unique_lock<std::mutex> lk(m);
    cv.wait(lk, [&] { return stop || job_id != seen; });
    if (stop) return;
    seen = job_id;
    const Job j = job;
    lk.unlock();
    // rows [r0, r1) of the job are split evenly over the threads
    const size_t per = (j.rows + nthreads - 1) / nthreads;
    const size_t r0 = std::min(j.rows, per * t), r1 = std::min(j.rows, r0 + per);
    if (j.src_pitch == j.row_words) {
      if (r1 > r0) memcpy(j.dst + r0 * j.row_words, j.src + r0 * j.src_pitch, (r1 - r0) * j.row_words * sizeof(u32));
    } else {
      for (size_t r = r0; r < r1; r++) memcpy(j.dst + r * j.row_words, j.src + r * j.src_pitch, j.row_words * sizeof(u32));
    }
    lk.lock();
    if (++done == nthreads) cv_done.notify_all();
  }
}
// dst (device, dense rows of row_words) <- src (pageable host, rows src_pitch words apart)
void HostStager::copy_2d(u32* dst, const u32* src, size_t row_words, size_t src_pitch, size_t rows, cudaStream_t s) {
  std::lock_guard<std::mutex> use(use_mu);     // one staged copy at a time (callers hold the copy lock anyway)
  ensure();
  const size_t rows_per_slot = std::max<size_t>(1, slot_bytes / (row_words * sizeof(u32)));
  if (row_words * sizeof(u32) > slot_bytes) throw std::runtime_error("zkb200: row wider than a staging slot");
  for (size_t r0 = 0; r0 < rows; r0 += rows_per_slot) {
    const size_t nr = std::min(rows_per_slot, rows - r0);
    const int slot = next_slot;
    next_slot = (next_slot + 1) % nslots;
    ZKB_CUDA(cudaEventSynchronize(free_ev[slot]));     // the DMA that last read this slot is done
    {
      std::unique_lock<std::mutex> lk(m);
      job = Job{slots[slot], src + r0 * src_pitch, row_words, src_pitch, nr};
      done = 0;
      job_id++;
      cv.notify_all();
      cv_done.wait(lk, [&] { return done == nthreads; });
    }
    ZKB_CUDA(cudaMemcpyAsync(dst + r0 * row_words, slots[slot], nr * row_words * sizeof(u32), cudaMemcpyHostToDevice, s));
    ZKB_CUDA(cudaEventRecord(free_ev[slot], s));
  }
}

// ---- TwoAdicFriPcs::commit --------------------------------------------------------------------
static void check_height(const MachineInfo& M, size_t height, unsigned& ln) {
  ln = log2_exact(height);
  if (((size_t)1 << ln) != height) throw std::runtime_error("zkb200: trace height is not a power of two");
  if (ln + M.log_blowup > 24) throw std::runtime_error("zkb200: LDE height exceeds the field's two-adicity (2^24)");
}

// Merkle tree over LDEs that are already in `out.ldes` (`pre`: row digests of height groups hashed
// piecewise during the upload); fetches the root.
static void pcs_merkle(Ctx& ctx, Lane& L, Commit& out, const std::map<unsigned, const u32*>* pre, const char* merkle_stage) {
  std::vector<MatRef> refs;
  for (size_t i = 0; i < out.ldes.size(); i++)
    refs.push_back(MatRef{out.ldes[i].d(), (u32)out.ldes[i].width, out.log_n[i] + ctx.machine.log_blowup});
  std::unique_ptr<StageTimer> tm(merkle_stage ? new StageTimer(ctx, L, merkle_stage) : nullptr);
  merkle_build(refs, L.arena, out.layers, L.d_small, L.stream, pre);
  ZKB_CUDA(cudaMemcpyAsync(L.h_small, L.d_small, 32, cudaMemcpyDeviceToHost, L.stream));
  ZKB_CUDA(cudaStreamSynchronize(L.stream));
  tm.reset();
  memcpy(out.root, L.h_small, 32);
  out.log_max_height = 0;
  for (auto& r : refs) out.log_max_height = std::max(out.log_max_height, r.log_height);
}

// coset LDE (shift GENERATOR / domain_shift) of every matrix + MMCS tree.
void pcs_commit(Ctx& ctx, Lane& L, std::vector<DevMat>& traces, const std::vector<Fp>& domain_shifts, Commit& out,
                const char* lde_stage, const char* merkle_stage) {
  const unsigned lb = ctx.machine.log_blowup;
  out.ldes.clear(); out.log_n.clear();
  Fp gen = fp_from_canonical(KB_GEN);
  {
    std::unique_ptr<StageTimer> tm(lde_stage ? new StageTimer(ctx, L, lde_stage) : nullptr);
    for (size_t i = 0; i < traces.size(); i++) {
      const DevMat& t = traces[i];
      unsigned ln;
      check_height(ctx.machine, t.height, ln);
      DevMat lde(t.height << lb, t.width, L.stream);
      Fp shift = gen * fp_inv(domain_shifts[i]);
      coset_lde_batch(ctx.tables, t.d(), t.height, lde.d(), lde.height, ln, t.width, lb, shift, L.stream, L.keep);
      out.ldes.push_back(std::move(lde));
      out.log_n.push_back(ln);
    }
  }
  pcs_merkle(ctx, L, out, nullptr, merkle_stage);
}

static void sort_traces(std::vector<TraceIn>& v) {
  std::stable_sort(v.begin(), v.end(), [](const TraceIn& a, const TraceIn& b) {
    if (a.height != b.height) return a.height > b.height;
    return a.name < b.name;
  });
}

Pk* prover_setup(Ctx& ctx, const std::vector<TraceIn>& prep_in, u32 pc_start, const u32* init_gsum) {
  ZKB_CUDA(cudaSetDevice(ctx.device));
  LaneGuard guard(ctx);
  Lane& L = *guard.lane;
  L.begin();
  std::unique_ptr<Pk> pk(new Pk());
  pk->ctx = &ctx;
  std::vector<TraceIn> prep = prep_in;
  sort_traces(prep);
  std::vector<Fp> shifts;
  for (auto& t : prep) {
    const ChipInfo* c = ctx.machine.find(t.name);
    if (!c) throw std::runtime_error("zkb200: setup: unknown chip " + t.name);
    if (c->prep_width != t.width) throw std::runtime_error("zkb200: setup: preprocessed width mismatch for " + t.name);
    pk->names.push_back(t.name);
    pk->local_only.push_back(c->local_only);
    pk->traces.push_back(upload_colmajor(ctx, t.data, t.height, t.width, L.stream, L.stream));
    shifts.push_back(fp_one());
  }
  if (!prep.empty()) {
    pcs_commit(ctx, L, pk->traces, shifts, pk->data, nullptr, nullptr);
    for (int i = 0; i < 8; i++) pk->commit_canon[i] = fp_to_canonical(fp_raw(pk->data.root[i]));
  }
  pk->pc_start = pc_start;
  for (int i = 0; i < 14; i++) pk->init_global_sum[i] = init_gsum ? init_gsum[i] : SEPTIC_DIGEST_ZERO[i];   // null: SepticDigest::zero(), as RecursionProgram (crates/recursion/core/src/runtime/program.rs:24-26)
  ZKB_CUDA(cudaStreamSynchronize(L.stream));
  return pk.release();
}

// ZKB200_TRACE_EVENTS: which row filler a chip has (csrc/tracegen.cu) and how long its event records are
static size_t event_record_words(const std::string& chip) {
  if (chip == "KeccakSponge") return KS_REC_WORDS;
  if (chip == "Global") return GLOBAL_EVENT_WORDS;
  if (alu_chip_by_name(chip.c_str()) >= 0) return (size_t)alu_event_words(alu_chip_by_name(chip.c_str()));
  throw std::runtime_error("zkb200: commit: no row filler for chip " + chip + " (ZKB200_TRACE_EVENTS)");
}
static void generate_trace_colmajor(const std::string& chip, const u32* events_dev, size_t n_events, size_t height, u32* out,
                                    cudaStream_t s) {
  if (chip == "KeccakSponge") keccak_sponge_trace(events_dev, n_events, height, out, s);
  else if (chip == "Global") global_trace(events_dev, n_events, height, out, true, s);
  else alu_trace(alu_chip_by_name(chip.c_str()), events_dev, n_events, height, out, true, s);
}

// CpuProver::commit (crates/stark/src/prover.rs:258-292), pipelined inside the shard.  Every matrix
// is cut into column PIECES of about ZKB200_PIECE_MB; a piece travels
//   host rows --(copy stream: 2-D DMA, or the pinned ring for pageable memory)--> row-major stage
//   --(compute lane)--> transpose into the column-major trace -> coset LDE of its columns
//   -> (a matrix alone in its height class) absorbed into the per-row sponge states,
// so the layout change, the LDE and the leaf hashing of piece k run under the upload of piece k+1 and
// one shard in flight (the reference's GPU options: shard_batch_size = 1, crates/stark/src/opts.rs:83-110)
// already overlaps PCIe with compute.  What is left after the last piece: the Merkle levels above the leaves.
Shard* prover_commit(Ctx& ctx, const std::vector<TraceIn>& traces_in, const u32* pv, size_t npv) {
  ZKB_CUDA(cudaSetDevice(ctx.device));
  nvtxRangePushA("commit main traces");
  struct PopRange { ~PopRange() { nvtxRangePop(); } } pop_range;
  std::unique_ptr<Shard> sh(new Shard());
  sh->ctx = &ctx;
  std::vector<TraceIn> traces = traces_in;
  sort_traces(traces);
  if (traces.empty()) throw std::runtime_error("zkb200: commit: no traces");
  const unsigned lb = ctx.machine.log_blowup;
  std::vector<unsigned> logn(traces.size());
  std::map<unsigned, int> group_size;
  for (size_t i = 0; i < traces.size(); i++) {
    const TraceIn& t = traces[i];
    const ChipInfo* c = ctx.machine.find(t.name);
    if (!c) throw std::runtime_error("zkb200: commit: unknown chip " + t.name);
    if (c->main_width != t.width) throw std::runtime_error("zkb200: commit: main width mismatch for " + t.name);
    check_height(ctx.machine, t.height, logn[i]);
    if ((t.flags & ~(TRACE_EVENTS | TRACE_COL_MAJOR)) || (t.flags & (TRACE_EVENTS | TRACE_COL_MAJOR)) == (TRACE_EVENTS | TRACE_COL_MAJOR))
      throw std::runtime_error("zkb200: commit: bad zkb200_trace.flags for " + t.name);
    if (t.flags & TRACE_EVENTS) {
      event_record_words(t.name);       // throws for a chip without a row filler
      const size_t rows_needed = t.name == "KeccakSponge" ? t.n_events * KS_ROUNDS : t.name == "Global" ? t.n_events
                                 : ceil_div(t.n_events, (size_t)alu_events_per_row(alu_chip_by_name(t.name.c_str())));
      const size_t w = t.name == "KeccakSponge" ? (size_t)KS_WIDTH : t.name == "Global" ? (size_t)GLOBAL_WIDTH
                       : (size_t)alu_width(alu_chip_by_name(t.name.c_str()));
      if (w != t.width) throw std::runtime_error("zkb200: commit: the row filler of " + t.name + " writes another width");
      if (rows_needed > t.height) throw std::runtime_error("zkb200: commit: more event rows than the table holds: " + t.name);
    } else if ((t.flags & TRACE_COL_MAJOR) && t.height * t.width && !is_device_pointer(t.data))
      throw std::runtime_error("zkb200: commit: ZKB200_TRACE_COL_MAJOR needs a device pointer: " + t.name);
    group_size[logn[i]]++;
  }
  struct Piece {
    size_t mat, col0, ncols;
    const u32* src = nullptr; size_t src_pitch = 0;     // row-major device source of the transpose
    DevBuf stage;
    std::shared_ptr<DevBuf> whole;                      // dma mode: the matrix's staging buffer, shared by its pieces
    cudaEvent_t ready = nullptr;
    int pull_index = -1; unsigned long long pull_tiles = 0;      // pull mode: counter and its final value
    bool last_of_matrix = false;
  };
  struct Events {      // RAII: an exception between creation and the end of the call must not leak them
    std::vector<cudaEvent_t> v;
    ~Events() { for (auto e : v) cudaEventDestroy(e); }
  } events;
  std::vector<Piece> pieces;
  const size_t piece_bytes = ctx.piece_bytes;
  for (size_t i = 0; i < traces.size(); i++) {
    const TraceIn& t = traces[i];
    if (t.height * t.width == 0) continue;
    size_t per = piece_bytes / (t.height * sizeof(u32));
    per = per >= t.width ? t.width : std::max<size_t>(8, per & ~(size_t)7);     // whole rate blocks (hash.cu: leaf_absorb)
    for (size_t c0 = 0; c0 < t.width; c0 += per) {
      Piece p;
      p.mat = i; p.col0 = c0; p.ncols = std::min(per, t.width - c0);
      p.last_of_matrix = c0 + p.ncols == t.width;
      pieces.push_back(std::move(p));
    }
  }
  // The heaviest height class crosses first: while the light ones follow, its LDE and leaf hashing already run
  // (the commitment does not depend on the order the CLASSES are processed in: LDEs land in their matrix's
  // slot and every class keeps its own sponge states; inside a class the pieces keep the commit order).
  std::map<unsigned, size_t> class_cells;
  for (size_t i = 0; i < traces.size(); i++) class_cells[logn[i]] += traces[i].height * traces[i].width;
  std::stable_sort(pieces.begin(), pieces.end(), [&](const Piece& a, const Piece& b) {
    const size_t sa = class_cells[logn[a.mat]], sb = class_cells[logn[b.mat]];
    if (sa != sb) return sa > sb;                                   // the heaviest height class first
    if (logn[a.mat] != logn[b.mat]) return logn[a.mat] > logn[b.mat];
    if (a.mat != b.mat) return a.mat < b.mat;                       // inside a class: commit order (the sponge's order)
    return a.col0 < b.col0;
  });
  // Phase 1 (copy stream, its own lock): the traces cross PCIe.  Other host threads may hold the compute
  // lanes meanwhile, the way the reference keeps several shards in flight
  // (crates/core/machine/src/utils/prove.rs:487-521).  Three kinds of source:
  //   device memory ......... nothing to copy; phase 2 transposes from it
  //   pinned host memory .... "pull": a few persistent CTAs on the (high-priority) copy stream read the
  //                           rows over PCIe and write the column-major trace directly (layout.cu);
  //                           ZKB200_UPLOAD=dma: one contiguous DMA per matrix into a staging buffer
  //   pageable host memory .. host threads gather the piece into the pinned ring, DMA, staging buffer
  for (size_t i = 0; i < traces.size(); i++) sh->names.push_back(traces[i].name);
  cudaEvent_t alloc_done = nullptr, pull_armed = nullptr, pull_finished = nullptr;
  std::vector<PullPiece> pull;
  std::map<size_t, DevBuf> event_bufs;      // ZKB200_TRACE_EVENTS: device copies of host event records, by matrix
  unsigned long long pull_tiles = 0;
  DevBuf pull_dev;
  u32* pull_done = nullptr;
  {
    std::lock_guard<std::mutex> lock(ctx.copy_mu);
    for (size_t i = 0; i < traces.size(); i++) sh->traces.push_back(DevMat(traces[i].height, traces[i].width, ctx.copy_stream));
    ZKB_CUDA(cudaEventCreateWithFlags(&alloc_done, cudaEventDisableTiming));
    events.v.push_back(alloc_done);
    ZKB_CUDA(cudaEventRecord(alloc_done, ctx.copy_stream));
    std::vector<std::shared_ptr<DevBuf>> whole(traces.size());     // dma mode: one staging buffer per matrix
    // if every matrix is a single piece there is nothing to pull: leave them all to the pull kernel then (it
    // needs no staging buffer) - the split only pays next to a multi-piece matrix
    bool any_multi = false;
    for (auto& p : pieces) if (!(p.col0 == 0 && p.last_of_matrix)) any_multi = true;
    const bool traces_multi_piece_only = !any_multi;
    for (auto& p : pieces) {
      const TraceIn& t = traces[p.mat];
      bool pinned = false;
      if (t.flags & (TRACE_EVENTS | TRACE_COL_MAJOR)) {
        // nothing crosses PCIe as a matrix: the lane generates the table from its events / copies the columns
        p.src = nullptr;
        if ((t.flags & TRACE_EVENTS) && p.col0 == 0 && t.n_events && !is_device_pointer(t.data, &pinned) && !pinned) {
          // pageable event records: the copy engine brings them over (a few MB), the lane waits for the event
          const size_t words = t.n_events * event_record_words(t.name);
          event_bufs[p.mat] = DevBuf(words, ctx.copy_stream);
          ZKB_CUDA(cudaMemcpyAsync(event_bufs[p.mat].p, t.data, words * sizeof(u32), cudaMemcpyHostToDevice, ctx.copy_stream));
          ZKB_CUDA(cudaEventCreateWithFlags(&p.ready, cudaEventDisableTiming));
          events.v.push_back(p.ready);
          ZKB_CUDA(cudaEventRecord(p.ready, ctx.copy_stream));
        }
        continue;
      }
      if (is_device_pointer(t.data, &pinned)) {
        p.src = t.data + p.col0; p.src_pitch = t.width;
        continue;
      }
      const u32* mapped = nullptr;
      if (pinned && ctx.upload_mode == UPLOAD_PULL && cudaHostGetDevicePointer((void**)&mapped, (void*)t.data, 0) != cudaSuccess) { cudaGetLastError(); mapped = nullptr; }
      // pull mode: a matrix that is ONE piece gains nothing from the column-wise pull; it goes by contiguous DMA
      // on a second copy stream, so the copy engine and the pull kernel share the link (the pull kernel alone
      // reaches ~40 GB/s next to the compute kernels, the link carries 55)
      const bool dma_side = mapped && ctx.wait_value && ctx.pull_split && p.col0 == 0 && p.last_of_matrix && !traces_multi_piece_only;
      if (dma_side) {
        // the staging buffer comes from the main copy stream (stream-ordered pool); the side stream starts after that
        p.stage = DevBuf(t.height * t.width, ctx.copy_stream);
        cudaEvent_t staged;
        ZKB_CUDA(cudaEventCreateWithFlags(&staged, cudaEventDisableTiming));
        events.v.push_back(staged);
        ZKB_CUDA(cudaEventRecord(staged, ctx.copy_stream));
        ZKB_CUDA(cudaStreamWaitEvent(ctx.dma_stream, staged, 0));
        ZKB_CUDA(cudaMemcpyAsync(p.stage.p, t.data, t.height * t.width * sizeof(u32), cudaMemcpyHostToDevice, ctx.dma_stream));
        p.src = p.stage.p; p.src_pitch = t.width;
        ZKB_CUDA(cudaEventCreateWithFlags(&p.ready, cudaEventDisableTiming));
        events.v.push_back(p.ready);
        ZKB_CUDA(cudaEventRecord(p.ready, ctx.dma_stream));
        continue;
      }
      if (mapped && ctx.wait_value) {
        // one persistent kernel pulls all such pieces (launched below); the lane waits for the piece's tile count
        PullPiece pp;
        const uintptr_t addr = (uintptr_t)mapped;
        pp.base = (const u32*)(addr & ~(uintptr_t)127);
        pp.word_off = (addr & 127) / sizeof(u32) + p.col0;
        pp.pitch = t.width;
        pp.dst = sh->traces[p.mat].d() + p.col0 * t.height;
        pp.rows = t.height; pp.cols = p.ncols;
        pp.col_tiles = pull_piece_col_tiles(p.ncols);
        pp.tile_begin = pull_tiles;
        p.pull_index = (int)pull.size();
        p.pull_tiles = pull_piece_tiles(t.height, p.ncols);
        pull_tiles += p.pull_tiles;
        pull.push_back(pp);
        p.src = nullptr;                                   // already column-major when its counter is full
        continue;
      } else if (pinned && ctx.upload_mode != UPLOAD_DMA) {
        // 2-D DMA of the column slice (ZKB200_UPLOAD=dma2d)
        p.stage = DevBuf(t.height * p.ncols, ctx.copy_stream);
        ZKB_CUDA(cudaMemcpy2DAsync(p.stage.p, p.ncols * sizeof(u32), t.data + p.col0, t.width * sizeof(u32), p.ncols * sizeof(u32),
                                   t.height, cudaMemcpyHostToDevice, ctx.copy_stream));
        p.src = p.stage.p; p.src_pitch = p.ncols;
      } else if (pinned) {
        if (!whole[p.mat]) {
          whole[p.mat] = std::make_shared<DevBuf>(t.height * t.width, ctx.copy_stream);
          ZKB_CUDA(cudaMemcpyAsync(whole[p.mat]->p, t.data, t.height * t.width * sizeof(u32), cudaMemcpyHostToDevice, ctx.copy_stream));
        }
        p.whole = whole[p.mat];
        p.src = whole[p.mat]->p + p.col0; p.src_pitch = t.width;
      } else {
        p.stage = DevBuf(t.height * p.ncols, ctx.copy_stream);
        ctx.stager.copy_2d(p.stage.p, t.data + p.col0, p.ncols, t.width, t.height, ctx.copy_stream);
        p.src = p.stage.p; p.src_pitch = p.ncols;
      }
      ZKB_CUDA(cudaEventCreateWithFlags(&p.ready, cudaEventDisableTiming));
      events.v.push_back(p.ready);
      ZKB_CUDA(cudaEventRecord(p.ready, ctx.copy_stream));
    }
    if (!pull.empty()) {
      pull_dev = DevBuf((pull.size() * sizeof(PullPiece) + 3) / 4, ctx.copy_stream);
      if (pull.size() > Ctx::PULL_COUNTER_RING / 4) throw std::runtime_error("zkb200: commit: too many pieces");
      if (ctx.pull_counter_next + pull.size() > Ctx::PULL_COUNTER_RING) ctx.pull_counter_next = 0;     // under copy_mu
      pull_done = ctx.pull_counters + ctx.pull_counter_next;
      ctx.pull_counter_next += pull.size();
      ZKB_CUDA(cudaMemcpyAsync(pull_dev.p, pull.data(), pull.size() * sizeof(PullPiece), cudaMemcpyHostToDevice, ctx.copy_stream));
      ZKB_CUDA(cudaMemsetAsync(pull_done, 0, pull.size() * sizeof(u32), ctx.copy_stream));
      // the lane must not look at the counters before they are cleared
      ZKB_CUDA(cudaEventCreateWithFlags(&pull_armed, cudaEventDisableTiming));
      events.v.push_back(pull_armed);
      ZKB_CUDA(cudaEventRecord(pull_armed, ctx.copy_stream));
      pull_shard(reinterpret_cast<const PullPiece*>(pull_dev.p), (int)pull.size(), pull_tiles, pull_done, ctx.pull_ctas, ctx.pull_exclusive, ctx.copy_stream);
      ZKB_CUDA(cudaEventCreateWithFlags(&pull_finished, cudaEventDisableTiming));
      events.v.push_back(pull_finished);
      ZKB_CUDA(cudaEventRecord(pull_finished, ctx.copy_stream));
    }
  }
  // Phase 2 (a compute lane): layout change, LDE and leaf hashing piece by piece, then the tree
  LaneGuard guard(ctx);
  Lane& L = *guard.lane;
  L.begin();
  if (ctx.profile) { std::lock_guard<std::mutex> lock(ctx.stage_mu); ctx.stage_ms.clear(); }
  Commit& out = sh->main;
  // the traces were allocated on the copy stream: order the lane after that, release them in lane order
  ZKB_CUDA(cudaStreamWaitEvent(L.stream, alloc_done, 0));
  if (pull_armed) ZKB_CUDA(cudaStreamWaitEvent(L.stream, pull_armed, 0));
  pull_dev.stream = L.stream;
  for (size_t i = 0; i < traces.size(); i++) {
    sh->traces[i].buf.stream = L.stream;
    out.ldes.push_back(DevMat(traces[i].height << lb, traces[i].width, L.stream));
    out.log_n.push_back(logn[i]);
  }
  const Fp shift = fp_from_canonical(KB_GEN);     // trace domains are the subgroups themselves
  // Leaf hashing under the upload, per height class: the columns of its matrices (commit order = the order
  // the pieces are processed in) are absorbed in whole rate blocks as soon as their LDE is queued; up to 7
  // columns wait for the next piece; the last piece of the class takes the tail and writes the digests.
  struct Group { size_t pieces_left = 0; bool started = false; std::vector<const u32*> pending; DevBuf state, digests; };
  std::map<unsigned, Group> groups;
  for (auto& p : pieces) groups[logn[p.mat]].pieces_left++;
  std::map<unsigned, const u32*> pre;
  for (auto& p : pieces) {
    const TraceIn& t = traces[p.mat];
    const size_t n = t.height, H = n << lb;
    u32* cols = sh->traces[p.mat].d() + p.col0 * n;
    u32* lde_cols = out.ldes[p.mat].d() + p.col0 * H;
    if ((t.flags & TRACE_EVENTS) && p.col0 == 0) {
      // MachineAir::generate_trace on the device: the whole table at once, column-major, in place
      StageTimer tm(ctx, L, "commit_main_generate_traces");
      if (p.ready) ZKB_CUDA(cudaStreamWaitEvent(L.stream, p.ready, 0));
      const u32* ev = t.data;
      bool pinned = false;
      auto eb = event_bufs.find(p.mat);
      if (eb != event_bufs.end()) { eb->second.stream = L.stream; ev = eb->second.p; }
      else if (t.n_events && !is_device_pointer(t.data, &pinned)) {
        const size_t words = t.n_events * event_record_words(t.name);
        DevBuf& b = event_bufs[p.mat] = DevBuf(words, L.stream);
        pull_words(b.p, t.data, words, L.stream);        // pinned records: SM loads, not the copy engine (common.h)
        ev = b.p;
      }
      generate_trace_colmajor(t.name, ev, t.n_events, n, sh->traces[p.mat].d(), L.stream);
    }
    {
      StageTimer tm(ctx, L, "commit_main_wait_upload_transpose");
      if (p.ready) ZKB_CUDA(cudaStreamWaitEvent(L.stream, p.ready, 0));
      if (p.pull_index >= 0) ctx.wait_value(L.stream, pull_done + p.pull_index, (u32)p.pull_tiles);
      if (p.src) transpose_piece_to_colmajor(p.src, p.src_pitch, cols, n, p.ncols, L.stream);
      else if (t.flags & TRACE_COL_MAJOR)
        ZKB_CUDA(cudaMemcpyAsync(cols, t.data + p.col0 * n, p.ncols * n * sizeof(u32), cudaMemcpyDeviceToDevice, L.stream));
      p.stage.stream = L.stream;     // released in lane order, after the transpose
      p.stage.release();
      if (p.whole) { p.whole->stream = L.stream; p.whole.reset(); }
    }
    {
      StageTimer tm(ctx, L, "commit_main_lde");
      coset_lde_batch(ctx.tables, cols, n, lde_cols, H, logn[p.mat], p.ncols, lb, shift, L.stream, L.keep);
    }
    {
      StageTimer tm(ctx, L, "commit_main_merkle");
      const unsigned lh = logn[p.mat] + lb;
      Group& g = groups[logn[p.mat]];
      const bool last = --g.pieces_left == 0;
      for (size_t c = 0; c < p.ncols; c++) g.pending.push_back(lde_cols + c * H);
      const size_t take = last ? g.pending.size() : (g.pending.size() & ~(size_t)7);
      if (take || last) {
        const bool first = !g.started;
        if (first && !last) g.state = DevBuf(16 * H, L.stream);
        if (last) g.digests = DevBuf(8 * H, L.stream);
        const u32* const* ptrs = L.arena.push(g.pending.data(), std::max<size_t>(take, 1));
        leaf_absorb(ptrs, H, (u32)take, g.state.p, first, last, g.digests.p, L.stream);
        g.pending.erase(g.pending.begin(), g.pending.begin() + take);
        g.started = true;
        if (last) { pre[lh] = g.digests.p; g.state.release(); }
      }
    }
  }
  if (pull_finished) ZKB_CUDA(cudaStreamWaitEvent(L.stream, pull_finished, 0));    // the work list and counters are released in lane order
  pcs_merkle(ctx, L, out, &pre, "commit_main_merkle");
  sh->public_values.assign(pv, pv + npv);
  return sh.release();
}

// ZKB200_TRACE_DERIVED: a table that only receives lookups (Byte, Program) is not handed over at all; its multiplicity columns
// are counted from the rows of the shard's other tables (csrc/derive.cuh) - in the reference the host accumulates them while
// it generates those tables (bytes/trace.rs:46-67, program/mod.rs:115-158).  Every other table is first made resident in its
// column-major form (row fillers for event records, the layout change for uploaded rows), then the derived tables are
// counted, then the ordinary commit runs over device-resident column-major tables.  An optional path: it holds the traces
// twice for the duration of the call and drains the device before it returns.
Shard* prover_commit_derived(Ctx& ctx, const Pk& pk, const std::vector<TraceIn>& traces_in, const u32* pv, size_t npv) {
  ZKB_CUDA(cudaSetDevice(ctx.device));
  std::vector<TraceIn> resident = traces_in;
  std::vector<DevMat> owned;
  std::vector<DevBuf> event_bufs;
  owned.reserve(resident.size());
  event_bufs.reserve(resident.size());
  struct Drain { ~Drain() { cudaDeviceSynchronize(); } } drain;     // before `owned` goes: the commit's copies read it (declared last = runs first)
  {
    LaneGuard guard(ctx);
    Lane& L = *guard.lane;
    L.begin();
    cudaStream_t s = L.stream;
    for (auto& t : resident) {
      const ChipInfo* c = ctx.machine.find(t.name);
      if (!c) throw std::runtime_error("zkb200: commit: unknown chip " + t.name);
      if (c->main_width != t.width) throw std::runtime_error("zkb200: commit: main width mismatch for " + t.name);
      if (t.flags & ~(TRACE_EVENTS | TRACE_COL_MAJOR | TRACE_DERIVED)) throw std::runtime_error("zkb200: commit: bad zkb200_trace.flags for " + t.name);
      if ((t.flags & TRACE_DERIVED) && t.flags != TRACE_DERIVED) throw std::runtime_error("zkb200: commit: ZKB200_TRACE_DERIVED excludes the other flags: " + t.name);
      if ((t.flags & TRACE_EVENTS) && (t.flags & TRACE_COL_MAJOR)) throw std::runtime_error("zkb200: commit: bad zkb200_trace.flags for " + t.name);
      if (t.flags & (TRACE_DERIVED | TRACE_COL_MAJOR)) continue;
      if (t.height * t.width == 0) continue;
      if (t.flags & TRACE_EVENTS) {
        const size_t rec = event_record_words(t.name);
        const size_t rows_needed = t.name == "KeccakSponge" ? t.n_events * KS_ROUNDS : t.name == "Global" ? t.n_events
                                   : ceil_div(t.n_events, (size_t)alu_events_per_row(alu_chip_by_name(t.name.c_str())));
        if (rows_needed > t.height) throw std::runtime_error("zkb200: commit: more event rows than the table holds: " + t.name);
        const u32* ev = t.data;
        if (t.n_events && !is_device_pointer(t.data)) {
          event_bufs.emplace_back(t.n_events * rec, s);
          ZKB_CUDA(cudaMemcpyAsync(event_bufs.back().p, t.data, t.n_events * rec * sizeof(u32), cudaMemcpyHostToDevice, s));
          ev = event_bufs.back().p;
        }
        owned.emplace_back(t.height, t.width, s);
        generate_trace_colmajor(t.name, ev, t.n_events, t.height, owned.back().d(), s);
      } else {
        owned.push_back(upload_colmajor(ctx, t.data, t.height, t.width, s, s));
      }
      t.data = owned.back().d();
      t.flags = TRACE_COL_MAJOR;
      t.n_events = 0;
    }
    std::vector<std::pair<size_t, const u32*>> derived;
    for (size_t i = 0; i < resident.size(); i++) {
      TraceIn& t = resident[i];
      if (!(t.flags & TRACE_DERIVED)) continue;
      const ChipInfo* r = ctx.machine.find(t.name);
      const int pi = pk.index_of(t.name);
      if (pi < 0 || pk.traces[pi].height != t.height) throw std::runtime_error("zkb200: commit: ZKB200_TRACE_DERIVED needs the table's preprocessed trace at this height: " + t.name);
      std::vector<DeriveSender> snd;
      for (auto& o : resident) {
        if ((o.flags & TRACE_DERIVED) || o.height * o.width == 0) continue;
        const int qi = pk.index_of(o.name);
        if (qi >= 0 && pk.traces[qi].height != o.height) throw std::runtime_error("zkb200: commit: preprocessed and main height differ for " + o.name);
        snd.push_back({ctx.machine.find(o.name), DeriveTable{qi >= 0 ? pk.traces[qi].d() : nullptr, o.data, o.height}});
      }
      owned.emplace_back(t.height, t.width, s);
      derive_multiplicities(ctx.machine, *r, pk.traces[pi].d(), t.height, snd, owned.back().d(), s);
      derived.push_back({i, owned.back().d()});
    }
    for (auto& d : derived) { resident[d.first].data = d.second; resident[d.first].flags = TRACE_COL_MAJOR; }
    ZKB_CUDA(cudaStreamSynchronize(s));
  }
  return prover_commit(ctx, resident, pv, npv);
}

// ---- proof packing ------------------------------------------------------------------------------
namespace {
struct Writer {
  std::vector<u32> w;
  void put(u32 x) { w.push_back(x); }
  void put_fp(Fp x) { w.push_back(fp_to_canonical(x)); }
  void put_ef(const Ef& e) { for (int i = 0; i < 4; i++) put_fp(e.c[i]); }
  void put_digest_monty(const u32* d) { for (int i = 0; i < 8; i++) put_fp(fp_raw(d[i])); }
  void str(const std::string& s) {
    put((u32)s.size());
    for (size_t i = 0; i < s.size(); i += 4) {
      u32 x = 0;
      for (size_t b = 0; b < 4 && i + b < s.size(); b++) x |= (u32)(unsigned char)s[i + b] << (8 * b);
      put(x);
    }
  }
  size_t reserve(size_t n) { size_t at = w.size(); w.resize(at + n, 0); return at; }
};

struct OpenMat {           // one matrix of one round in the opening
  const DevMat* lde;
  unsigned log_h;          // LDE log height
  unsigned log_n;
  int npoints;             // 1: zeta; 2: zeta and zeta * g_n
  size_t ys_off;           // word offset of [npoints][W] EF in the opened-values buffer
  size_t alpha_off;        // num_reduced[log_h] before this matrix
};
}  // namespace

std::vector<u32> prover_open(Ctx& ctx, const Pk& pk, Shard& sh, u32* challenger34) {
  ZKB_CUDA(cudaSetDevice(ctx.device));
  LaneGuard guard(ctx);
  Lane& L = *guard.lane;
  L.begin();
  cudaStream_t s = L.stream;
  const MachineInfo& M = ctx.machine;
  const unsigned lb = M.log_blowup;
  const size_t nc = sh.names.size();
  std::vector<const ChipInfo*> chips;
  std::vector<unsigned> logn;
  for (size_t i = 0; i < nc; i++) { chips.push_back(M.find(sh.names[i])); logn.push_back(log2_exact(sh.traces[i].height)); }
  if (sh.public_values.size() < M.num_pv_elts) throw std::runtime_error("zkb200: open: fewer public values than num_pv_elts");
  for (auto& name : pk.names) if (std::find(sh.names.begin(), sh.names.end(), name) == sh.names.end())
    throw std::runtime_error("zkb200: open: preprocessed chip " + name + " missing from shard");

  Challenger ch;
  ch.load(challenger34);
  for (u32 i = 0; i < M.num_pv_elts; i++) ch.observe_canonical(sh.public_values[i]);
  ch.observe_digest(sh.main.root);
  const Ef perm_alpha = ch.sample_ext(), perm_beta = ch.sample_ext();

  // public values on the device (Montgomery)
  std::vector<u32> pv_m(std::max<size_t>(sh.public_values.size(), 1), 0);
  for (size_t i = 0; i < sh.public_values.size(); i++) pv_m[i] = fp_from_canonical(sh.public_values[i] % KB_P).v;
  const u32* pv_dev = L.arena.push(pv_m.data(), pv_m.size());

  // ---- permutation traces (K5) + commit -----------------------------------------------------
  std::vector<DevMat> perm_traces;
  std::vector<Ef> local_sums(nc);
  std::vector<std::array<u32, 14>> global_sums(nc);   // Montgomery
  {
    StageTimer tm(ctx, L, "permutation_trace");
    // sums land in d_small: per chip 4 words local + 14 words global
    if (nc > 200) throw std::runtime_error("zkb200: too many chips in shard");
    std::vector<GatherJob> jobs;
    for (size_t i = 0; i < nc; i++) {
      int pi = pk.index_of(sh.names[i]);
      const DevMat* prep = pi >= 0 ? &pk.traces[pi] : nullptr;
      if (prep && prep->height != sh.traces[i].height) throw std::runtime_error("zkb200: preprocessed and main have different heights: " + sh.names[i]);
      const size_t n = sh.traces[i].height;
      DevMat pt(n, 4 * chips[i]->perm_width_ef(), s);
      u32* sums = L.d_small + 4096 + i * 4;   // Montgomery scratch, converted by the gather below
      permutation_trace(M, *chips[i], prep ? prep->d() : nullptr, sh.traces[i].d(), n, perm_alpha, perm_beta, pt.d(), sums, s);
      jobs.push_back(GatherJob{sums, 1, 4, (u32)(i * 18)});
      if (chips[i]->global_scope)
        jobs.push_back(GatherJob{sh.traces[i].d() + (sh.traces[i].width - 14) * n + (n - 1), n, 14, (u32)(i * 18 + 4)});
      perm_traces.push_back(std::move(pt));
    }
    ZKB_CUDA(cudaMemsetAsync(L.d_small, 0, nc * 18 * 4, s));
    const GatherJob* jd = L.arena.push(jobs.data(), jobs.size());
    gather_canonical(jd, jobs.size(), L.d_small, s);
    ZKB_CUDA(cudaMemcpyAsync(L.h_small, L.d_small, nc * 18 * 4, cudaMemcpyDeviceToHost, s));
    ZKB_CUDA(cudaStreamSynchronize(s));
    for (size_t i = 0; i < nc; i++) {
      const u32* w = L.h_small + i * 18;   // canonical
      for (int c = 0; c < 4; c++) local_sums[i].c[c] = fp_from_canonical(w[c]);
      // a Local-scope chip carries SepticDigest::zero(), which is the curve's START point and not
      // 14 zeros (crates/stark/src/septic_digest.rs:9-42, prover.rs:352; checked by verifier.rs:102)
      for (int k = 0; k < 14; k++)
        global_sums[i][k] = fp_from_canonical(chips[i]->global_scope ? w[4 + k] : SEPTIC_DIGEST_ZERO[k]).v;
    }
  }
  Commit perm_commit;
  {
    std::vector<Fp> shifts(nc, fp_one());
    pcs_commit(ctx, L, perm_traces, shifts, perm_commit, "commit_permutation_lde", "commit_permutation_merkle");
  }
  perm_traces.clear();
  ch.observe_digest(perm_commit.root);
  for (size_t i = 0; i < nc; i++) {
    ch.observe_ext(local_sums[i]);
    for (int k = 0; k < 14; k++) ch.observe(fp_raw(global_sums[i][k]));
  }

  // ---- quotient (K3) + commit ---------------------------------------------------------------
  const Ef alpha = ch.sample_ext();
  std::vector<DevMat> quot_chunks;
  std::vector<Fp> quot_shifts;
  {
    StageTimer tm(ctx, L, "quotient");
    for (size_t i = 0; i < nc; i++) {
      const unsigned lqd = chips[i]->log_quotient_degree;
      const size_t n = (size_t)1 << logn[i], nchunks = (size_t)1 << lqd;
      DevBuf q(nchunks * 4 * n, s);
      QuotientInputs in;
      int pi = pk.index_of(sh.names[i]);
      in.prep_lde = pi >= 0 ? pk.data.ldes[pi].d() : nullptr;
      in.main_lde = sh.main.ldes[i].d();
      in.perm_lde = perm_commit.ldes[i].d();
      in.lde_h = n << lb;
      in.log_n = logn[i];
      in.perm_alpha = perm_alpha; in.perm_beta = perm_beta; in.local_sum = local_sums[i]; in.alpha = alpha;
      memcpy(in.global_sum, global_sums[i].data(), sizeof(in.global_sum));
      in.pub_dev = pv_dev;
      quotient_values(M, *chips[i], ctx.tables, in, q.p, s);
      Fp gq = two_adic_generator(logn[i] + lqd);
      for (size_t j = 0; j < nchunks; j++) {
        DevMat cm(n, 4, s);
        ZKB_CUDA(cudaMemcpyAsync(cm.d(), q.p + j * 4 * n, 4 * n * sizeof(u32), cudaMemcpyDeviceToDevice, s));
        quot_chunks.push_back(std::move(cm));
        quot_shifts.push_back(fp_from_canonical(KB_GEN) * fp_pow(gq, j));
      }
    }
  }
  Commit quot_commit;
  {
    pcs_commit(ctx, L, quot_chunks, quot_shifts, quot_commit, "commit_quotient_lde", "commit_quotient_merkle");
  }
  quot_chunks.clear();
  ch.observe_digest(quot_commit.root);
  const Ef zeta = ch.sample_ext();

  // ---- opening (K4a/K4b) ----------------------------------------------------------------------
  struct Round { const Commit* c; std::vector<OpenMat> mats; };
  std::vector<Round> rounds;
  size_t ys_words = 0;
  size_t num_reduced[32] = {0};
  size_t max_w = 1;
  auto add_round = [&](const Commit* c, const std::vector<int>& npts) {
    Round r; r.c = c;
    for (size_t i = 0; i < c->ldes.size(); i++) {
      OpenMat m;
      m.lde = &c->ldes[i]; m.log_n = c->log_n[i]; m.log_h = c->log_n[i] + lb; m.npoints = npts[i];
      m.ys_off = ys_words; ys_words += (size_t)m.npoints * m.lde->width * 4;
      m.alpha_off = num_reduced[m.log_h];
      num_reduced[m.log_h] += (size_t)m.npoints * m.lde->width;
      max_w = std::max(max_w, m.lde->width);
      r.mats.push_back(m);
    }
    rounds.push_back(std::move(r));
  };
  if (!pk.data.empty()) { std::vector<int> np; for (size_t i = 0; i < pk.names.size(); i++) np.push_back(pk.local_only[i] ? 1 : 2); add_round(&pk.data, np); }
  { std::vector<int> np; for (size_t i = 0; i < nc; i++) np.push_back(chips[i]->local_only ? 1 : 2); add_round(&sh.main, np); }
  add_round(&perm_commit, std::vector<int>(nc, 2));
  add_round(&quot_commit, std::vector<int>(quot_commit.ldes.size(), 1));
  const size_t main_round = pk.data.empty() ? 0 : 1;

  const Ef alpha_fri = ch.sample_ext();
  unsigned log_gmax = 0;
  for (auto& r : rounds) log_gmax = std::max(log_gmax, r.c->log_max_height);

  DevBuf ys_dev(std::max<size_t>(ys_words, 1), s);
  std::vector<DevBuf> ro(32);
  {
    StageTimer tm(ctx, L, "open_reduce");
    DevBuf apow(4 * max_w, s);
    ef_powers(alpha_fri, max_w, apow.p, s);
    // group by LDE height so that barycentric weights and inverse denominators are built once
    for (unsigned lh = lb; lh <= log_gmax; lh++) {
      bool any = false, any2 = false;
      for (auto& r : rounds) for (auto& m : r.mats) if (m.log_h == lh) { any = true; if (m.npoints == 2) any2 = true; }
      if (!any) continue;
      const unsigned ln = lh - lb;
      const size_t n = (size_t)1 << ln, H = (size_t)1 << lh;
      const Ef znext = zeta * two_adic_generator(ln);
      DevBuf w0(4 * n, s), w1(any2 ? 4 * n : 0, s), d0(4 * H, s), d1(any2 ? 4 * H : 0, s);
      bary_weights(ctx.tables, ln, zeta, w0.p, s);
      inv_denominators(ctx.tables, lh, zeta, d0.p, s);
      if (any2) { bary_weights(ctx.tables, ln, znext, w1.p, s); inv_denominators(ctx.tables, lh, znext, d1.p, s); }
      ro[lh] = DevBuf(4 * H, s);
      ZKB_CUDA(cudaMemsetAsync(ro[lh].p, 0, 4 * H * sizeof(u32), s));
      for (auto& r : rounds) for (auto& m : r.mats) {
        if (m.log_h != lh || m.lde->width == 0) continue;
        const size_t W = m.lde->width;
        u32* ys = ys_dev.p + m.ys_off;
        eval_columns(m.lde->d(), H, n, W, w0.p, w1.p, m.npoints, ys, s);
        Ef off0 = ef_pow(alpha_fri, m.alpha_off), off1 = ef_pow(alpha_fri, m.alpha_off + W);
        reduce_matrix(m.lde->d(), H, W, apow.p, ys, m.npoints, off0, off1, d0.p, d1.p, ro[lh].p, s);
      }
    }
  }

  // ---- FRI commit phase (K4c) -------------------------------------------------------------------
  struct FriLayer { DevBuf folded; size_t m; DigestLayers tree; u32 root[8]; };
  std::vector<FriLayer> fri_layers;
  Ef final_poly;
  {
    StageTimer tm(ctx, L, "fri_commit_phase");
    DevBuf cur = std::move(ro[log_gmax]);
    size_t m = (size_t)1 << log_gmax;
    const size_t blowup = (size_t)1 << lb;
    // The transcript of the commit phase runs ON THE DEVICE (hash.h, DevChallenger): per layer a one-thread kernel
    // observes the layer's root and samples beta where the fold kernel reads it, so the ~20 layers are enqueued
    // back to back and the host reads the roots, the final polynomial and the challenger once (it used to wait
    // for a 32-byte copy per layer).  tr: [0, 64) challenger image, then 8 words of root and 4 of beta per layer,
    // then the final evaluations.
    const size_t nlayers = log_gmax > lb ? log_gmax - lb : 0;
    const size_t tr_roots = 64, tr_final = tr_roots + 12 * nlayers, tr_words = tr_final + 4 * blowup;
    if (tr_words * sizeof(u32) > (1 << 16)) throw std::runtime_error("zkb200: FRI transcript exceeds the lane's scratch");
    DevBuf tr(tr_words, s);
    DevChallenger* dch = reinterpret_cast<DevChallenger*>(tr.p);
    static_assert(sizeof(DevChallenger) <= 64 * sizeof(u32), "challenger image fits its slot");
    challenger_to_device(ch, dch, s);
    size_t li = 0;
    while (m > blowup) {
      FriLayer FL;
      FL.m = m;
      u32* root_dev = tr.p + tr_roots + 12 * li;
      fri_commit_layer(cur.p, m, FL.tree, root_dev, s);
      challenger_observe_digest_sample_ext(dch, root_dev, root_dev + 8, s);
      const unsigned lnext = log2_exact(m) - 1;
      DevBuf next(4 * (m >> 1), s);
      fri_fold_dev_beta(ctx.tables, cur.p, m, root_dev + 8, ro[lnext].p, next.p, s);
      FL.folded = std::move(cur);
      cur = std::move(next);
      fri_layers.push_back(std::move(FL));
      m >>= 1;
      li++;
    }
    // `cur` holds blowup evaluations of a constant polynomial
    ZKB_CUDA(cudaMemcpyAsync(tr.p + tr_final, cur.p, 4 * m * sizeof(u32), cudaMemcpyDeviceToDevice, s));
    ZKB_CUDA(cudaMemcpyAsync(L.h_small, tr.p, tr_words * sizeof(u32), cudaMemcpyDeviceToHost, s));
    ZKB_CUDA(cudaStreamSynchronize(s));
    ch.load_device_image(*reinterpret_cast<const DevChallenger*>(L.h_small));
    for (size_t l = 0; l < fri_layers.size(); l++) memcpy(fri_layers[l].root, L.h_small + tr_roots + 12 * l, 32);
    const u32* fin = L.h_small + tr_final;
    for (int c = 0; c < 4; c++) final_poly.c[c] = fp_raw(fin[c * m]);
    for (size_t i = 1; i < m; i++)
      for (int c = 0; c < 4; c++)
        if (fin[c * m + i] != final_poly.c[c].v) throw std::runtime_error("zkb200: FRI final polynomial is not constant (unsatisfied constraints?)");
  }
  ch.observe_ext(final_poly);

  // ---- proof of work (K4d) ------------------------------------------------------------------
  u32 pow_witness = 0;
  {
    StageTimer tm(ctx, L, "grind");
    u32 st[16];
    for (int i = 0; i < 16; i++) st[i] = ch.state[i].v;
    for (unsigned i = 0; i < ch.n_in; i++) st[i] = ch.in_buf[i].v;
    pow_witness = grind_witness(st, ch.n_in, M.pow_bits, L.d_small + 8192, s);
    ch.observe_canonical(pow_witness);
    if (ch.sample_bits(M.pow_bits) != 0) throw std::runtime_error("zkb200: grinding produced an invalid witness");
  }

  // ---- assemble the proof; query openings are gathered on the device (K4e) ----------------------
  std::vector<u32> ys_host(ys_words);
  if (ys_words) ZKB_CUDA(cudaMemcpyAsync(ys_host.data(), ys_dev.p, ys_words * sizeof(u32), cudaMemcpyDeviceToHost, s));
  ZKB_CUDA(cudaStreamSynchronize(s));
  auto ys_at = [&](const OpenMat& m, int pt, size_t col) {
    const u32* p = ys_host.data() + m.ys_off + ((size_t)pt * m.lde->width + col) * 4;
    Ef e; for (int c = 0; c < 4; c++) e.c[c] = fp_raw(p[c]);
    return e;
  };

  Writer o;
  o.put(0x46504b5au); o.put(1);
  o.put_digest_monty(sh.main.root); o.put_digest_monty(perm_commit.root); o.put_digest_monty(quot_commit.root);
  o.put((u32)nc);
  size_t qidx = 0;
  for (size_t i = 0; i < nc; i++) {
    const ChipInfo& c = *chips[i];
    o.str(c.name);
    o.put(logn[i]);
    o.put(c.prep_width); o.put(c.main_width); o.put(4 * c.perm_width_ef()); o.put(1u << c.log_quotient_degree);
    auto put_opened = [&](const OpenMat* m, size_t width) {
      for (int pt = 0; pt < 2; pt++)
        for (size_t col = 0; col < width; col++) {
          if (m && pt < m->npoints) o.put_ef(ys_at(*m, pt, col));
          else o.put_ef(ef_zero());
        }
    };
    int pi = pk.index_of(c.name);
    put_opened(pi >= 0 ? &rounds[0].mats[pi] : nullptr, pi >= 0 ? c.prep_width : 0);
    put_opened(&rounds[main_round].mats[i], c.main_width);
    put_opened(&rounds[main_round + 1].mats[i], 4 * c.perm_width_ef());
    for (u32 j = 0; j < (1u << c.log_quotient_degree); j++, qidx++)
      for (size_t col = 0; col < 4; col++) o.put_ef(ys_at(rounds[main_round + 2].mats[qidx], 0, col));
    for (int k = 0; k < 14; k++) o.put_fp(fp_raw(global_sums[i][k]));
    o.put_ef(local_sums[i]);
  }
  o.put((u32)sh.public_values.size());
  for (u32 x : sh.public_values) o.put(x);
  o.put((u32)fri_layers.size());
  for (auto& fl : fri_layers) o.put_digest_monty(fl.root);
  o.put_ef(final_poly);
  o.put(pow_witness);
  o.put(M.num_queries);

  std::vector<GatherJob> jobs;
  {
    StageTimer tm(ctx, L, "query_openings");
    for (u32 q = 0; q < M.num_queries; q++) {
      const size_t index = ch.sample_bits(log_gmax);
      o.put((u32)rounds.size());
      for (auto& r : rounds) {
        const unsigned lmax = r.c->log_max_height;
        const size_t ridx = index >> (log_gmax - lmax);
        o.put((u32)r.mats.size());
        for (auto& m : r.mats) {
          const size_t W = m.lde->width, H = m.lde->height;
          const size_t row = ridx >> (lmax - m.log_h);
          o.put((u32)W);
          size_t at = o.reserve(W);
          if (W) jobs.push_back(GatherJob{m.lde->d() + row, H, (u32)W, (u32)at});
        }
        o.put(lmax);
        for (unsigned l = 0; l < lmax; l++) {
          size_t at = o.reserve(8);
          size_t node = (ridx >> l) ^ 1;
          jobs.push_back(GatherJob{r.c->layers.layer(l) + node, r.c->layers.count[l], 8, (u32)at});
        }
      }
      o.put((u32)fri_layers.size());
      for (size_t li = 0; li < fri_layers.size(); li++) {
        const FriLayer& FL = fri_layers[li];
        const size_t idx_i = index >> li, pair = idx_i >> 1, sib = idx_i ^ 1;
        size_t at = o.reserve(4);
        jobs.push_back(GatherJob{FL.folded.p + sib, FL.m, 4, (u32)at});
        const unsigned depth = (unsigned)FL.tree.count.size() - 1;
        o.put(depth);
        for (unsigned l = 0; l < depth; l++) {
          size_t a2 = o.reserve(8);
          size_t node = (pair >> l) ^ 1;
          jobs.push_back(GatherJob{FL.tree.layer(l) + node, FL.tree.count[l], 8, (u32)a2});
        }
      }
    }
    if (o.w.size() >= (1ull << 32)) throw std::runtime_error("zkb200: proof too large");
    // gather straight into a device image of the proof, then overlay the gathered words
    DevBuf jobs_dev((jobs.size() * sizeof(GatherJob) + 3) / 4 + 1, s);
    DevBuf img(o.w.size(), s);
    // the job list and the proof skeleton go up through the lane's pinned buffer, pulled by a kernel
    // (not the copy engine: common.h, pull_words)
    const size_t job_words = (jobs.size() * sizeof(GatherJob) + 3) / 4;
    if ((job_words + o.w.size()) * sizeof(u32) > L.h_big_bytes) throw std::runtime_error("zkb200: proof skeleton exceeds the lane's staging buffer");
    memcpy(L.h_big, jobs.data(), jobs.size() * sizeof(GatherJob));
    memcpy(L.h_big + job_words, o.w.data(), o.w.size() * sizeof(u32));
    pull_words(jobs_dev.p, L.h_big, job_words, s);
    pull_words(img.p, L.h_big + job_words, o.w.size(), s);
    gather_canonical(reinterpret_cast<const GatherJob*>(jobs_dev.p), jobs.size(), img.p, s);
    ZKB_CUDA(cudaMemcpyAsync(o.w.data(), img.p, o.w.size() * sizeof(u32), cudaMemcpyDeviceToHost, s));
    ZKB_CUDA(cudaStreamSynchronize(s));
  }
  ch.store(challenger34);
  return std::move(o.w);
}

}  // namespace zkb

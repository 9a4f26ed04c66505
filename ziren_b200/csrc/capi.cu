// extern "C" surface of libzkb200.so — see include/zkb200.h for the contract.
#include <cstring>
#include "../../include/zkb200.h"
#include <cstdlib>
#include <memory>
#include "derive.h"
#include "layout.h"
#include "logup.h"
#include "open.h"
#include "prover.h"
#include "quotient.h"
#include "quotient_codegen.h"
#include "tracegen.h"

using namespace zkb;
namespace zkb { extern std::atomic<int> g_ntt_force_k2; extern std::atomic<int> g_eval_v2; extern std::atomic<int> g_ntt_lean; extern std::atomic<int> g_quotient_codegen; extern std::atomic<int> g_k4b_rows; extern std::atomic<int> g_qk_block; extern std::atomic<int> g_logup_codegen;
               extern std::atomic<unsigned long long> g_quotient_generated_launches, g_quotient_interpreter_launches; }

// One prover object over one or several GPUs.  `c` (= *devs[0]) serves the kernel-level entry points;
// commit routes every shard to the device with the fewest shards in flight, open follows the shard,
// the proving key is replicated per device at setup - the object prove_with_context's worker threads
// can share (crates/core/machine/src/utils/prove.rs:487-521 calls ONE prover from all of them).
struct zkb200_ctx {
  std::vector<std::unique_ptr<Ctx>> devs;
  std::unique_ptr<std::atomic<int>[]> inflight;
  Ctx& c;
  explicit zkb200_ctx(std::vector<std::unique_ptr<Ctx>>&& d)
      : devs(std::move(d)), inflight(new std::atomic<int>[devs.size()]), c(*devs[0]) {
    for (size_t i = 0; i < devs.size(); i++) inflight[i] = 0;
  }
  int index_of_device(int device) const {
    for (size_t i = 0; i < devs.size(); i++) if (devs[i]->device == device) return (int)i;
    return -1;
  }
  // the device that already holds the (device-resident) traces, else the least loaded one
  int pick(const zkb200_trace* t, int n) const {
    for (int i = 0; i < n; i++) {
      cudaPointerAttributes attr;
      if (t[i].data && cudaPointerGetAttributes(&attr, t[i].data) == cudaSuccess && attr.type == cudaMemoryTypeDevice) {
        int k = index_of_device(attr.device);
        if (k >= 0) return k;
      } else cudaGetLastError();
    }
    int best = 0;
    for (size_t i = 1; i < devs.size(); i++) if (inflight[i].load() < inflight[best].load()) best = (int)i;
    return best;
  }
};
struct zkb200_pk { std::vector<Pk*> per_dev; Pk* p; };      // p = per_dev[0]
struct zkb200_shard { Shard* s; zkb200_ctx* owner; int dev; };

// MachineProver::Error text of the last failed call made by THIS host thread
static thread_local std::string g_err;

template <class F>
static int guarded(zkb200_ctx* ctx, F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    cudaGetLastError();
    // error path only: kernels of the failed call may still be queued on a lane and read buffers the
    // caller is about to release (stream-ordered frees on another stream) - drain the device first
    if (ctx) for (auto& d : ctx->devs) { cudaSetDevice(d->device); cudaDeviceSynchronize(); cudaGetLastError(); }
    return 1;
  }
}

static std::vector<TraceIn> to_traces(const zkb200_trace* t, int n) {
  std::vector<TraceIn> v;
  for (int i = 0; i < n; i++) v.push_back(TraceIn{t[i].name, t[i].data, t[i].height, t[i].width, t[i].flags, t[i].n_events});
  return v;
}
static Ef ef_from_canon(const uint32_t* w) { Ef e; for (int i = 0; i < 4; i++) e.c[i] = fp_from_canonical(w[i]); return e; }

extern "C" {

int zkb200_ctx_create_multi(const int* devices, int n_devices, const uint32_t* desc, size_t n_words, zkb200_ctx** out) {
  *out = nullptr;
  std::vector<std::unique_ptr<Ctx>> devs;
  int rc = guarded(nullptr, [&] {
    std::vector<int> ids;
    if (n_devices <= 0) {
      int count = 0;
      ZKB_CUDA(cudaGetDeviceCount(&count));
      for (int i = 0; i < count; i++) ids.push_back(i);
    } else ids.assign(devices, devices + n_devices);
    if (ids.empty()) throw std::runtime_error("zkb200: no CUDA device");
    for (int id : ids) {
      devs.emplace_back(new Ctx());
      devs.back()->init(id, desc, n_words);
    }
  });
  if (rc) {
    for (auto& d : devs) d->destroy();
    return rc;
  }
  *out = new zkb200_ctx(std::move(devs));
  return 0;
}
int zkb200_ctx_create(int device, const uint32_t* desc, size_t n_words, zkb200_ctx** out) {
  if (device < 0) return zkb200_ctx_create_multi(nullptr, 0, desc, n_words, out);     // every visible GPU
  return zkb200_ctx_create_multi(&device, 1, desc, n_words, out);
}
int zkb200_ctx_num_devices(const zkb200_ctx* ctx) { return (int)ctx->devs.size(); }
void zkb200_ctx_destroy(zkb200_ctx* ctx) {
  if (!ctx) return;
  for (auto& d : ctx->devs) d->destroy();
  delete ctx;
}
const char* zkb200_last_error(zkb200_ctx* ctx) { (void)ctx; return g_err.c_str(); }
void* zkb200_ctx_stream(zkb200_ctx* ctx) { return (void*)ctx->c.lanes[0].stream; }

int zkb200_setup(zkb200_ctx* ctx, const zkb200_trace* prep, int n, uint32_t pc_start, const uint32_t* init_global_sum,
                 uint32_t commit_out[8], zkb200_pk** out) {
  return guarded(ctx, [&] {
    std::unique_ptr<zkb200_pk> pk(new zkb200_pk{{}, nullptr});
    try {
      for (auto& d : ctx->devs) pk->per_dev.push_back(prover_setup(*d, to_traces(prep, n), pc_start, init_global_sum));
    } catch (...) {
      for (Pk* q : pk->per_dev) { cudaSetDevice(q->ctx->device); delete q; }
      throw;
    }
    pk->p = pk->per_dev[0];
    for (Pk* q : pk->per_dev)
      if (memcmp(q->commit_canon, pk->p->commit_canon, 32) != 0) throw std::runtime_error("zkb200: setup: devices disagree on the preprocessed commitment");
    if (commit_out) memcpy(commit_out, pk->p->commit_canon, 32);
    *out = pk.release();
  });
}
void zkb200_pk_free(zkb200_pk* pk) {
  if (!pk) return;
  for (Pk* q : pk->per_dev) { cudaSetDevice(q->ctx->device); delete q; }
  delete pk;
}
int zkb200_pk_initial_challenger(const zkb200_pk* pk, uint32_t challenger[34]) {
  Challenger ch;
  for (int i = 0; i < 8; i++) ch.observe_canonical(pk->p->commit_canon[i]);
  ch.observe_canonical(pk->p->pc_start);
  for (int i = 0; i < 14; i++) ch.observe_canonical(pk->p->init_global_sum[i]);
  ch.observe(fp_zero());
  ch.store(challenger);
  return 0;
}

int zkb200_commit(zkb200_ctx* ctx, const zkb200_trace* traces, int n, const uint32_t* pv, size_t npv, uint32_t commit_out[8],
                  zkb200_shard** out) {
  return guarded(ctx, [&] {
    const int k = ctx->pick(traces, n);
    ctx->inflight[k]++;
    Shard* s = nullptr;
    try {
      s = prover_commit(*ctx->devs[k], to_traces(traces, n), pv, npv);
    } catch (...) {
      ctx->inflight[k]--;
      throw;
    }
    if (commit_out) for (int i = 0; i < 8; i++) commit_out[i] = fp_to_canonical(fp_raw(s->main.root[i]));
    *out = new zkb200_shard{s, ctx, k};
  });
}
int zkb200_shard_device(const zkb200_shard* sh) { return sh->s->ctx->device; }
void zkb200_shard_free(zkb200_shard* sh) {
  if (!sh) return;
  if (sh->s) { cudaSetDevice(sh->s->ctx->device); delete sh->s; }
  if (sh->owner) sh->owner->inflight[sh->dev]--;
  delete sh;
}

int zkb200_open(zkb200_ctx* ctx, const zkb200_pk* pk, zkb200_shard* shard, uint32_t challenger[34], uint32_t** proof_words,
                size_t* n_words) {
  return guarded(ctx, [&] {
    if (shard->owner != ctx || pk->per_dev.size() != ctx->devs.size()) throw std::runtime_error("zkb200: open: shard / proving key belong to another context");
    std::vector<u32> w = prover_open(*ctx->devs[shard->dev], *pk->per_dev[shard->dev], *shard->s, challenger);
    *proof_words = (uint32_t*)malloc(w.size() * 4);
    memcpy(*proof_words, w.data(), w.size() * 4);
    *n_words = w.size();
  });
}
int zkb200_prove_shard(zkb200_ctx* ctx, const zkb200_pk* pk, const zkb200_trace* traces, int n, const uint32_t* pv, size_t npv,
                       uint32_t challenger[34], uint32_t** proof_words, size_t* n_words) {
  zkb200_shard* sh = nullptr;
  bool derived = false;
  for (int i = 0; i < n; i++) derived = derived || (traces[i].flags & TRACE_DERIVED);
  int rc;
  if (derived) {       // ZKB200_TRACE_DERIVED: the commit needs the proving key's preprocessed tables (prover.cu: prover_commit_derived)
    rc = guarded(ctx, [&] {
      if (pk->per_dev.size() != ctx->devs.size()) throw std::runtime_error("zkb200: prove_shard: the proving key belongs to another context");
      const int k = ctx->pick(traces, n);
      ctx->inflight[k]++;
      Shard* s = nullptr;
      try {
        s = prover_commit_derived(*ctx->devs[k], *pk->per_dev[k], to_traces(traces, n), pv, npv);
      } catch (...) {
        ctx->inflight[k]--;
        throw;
      }
      sh = new zkb200_shard{s, ctx, k};
    });
  } else rc = zkb200_commit(ctx, traces, n, pv, npv, nullptr, &sh);
  if (rc) return rc;
  rc = zkb200_open(ctx, pk, sh, challenger, proof_words, n_words);
  zkb200_shard_free(sh);
  return rc;
}
void zkb200_free(void* p) { free(p); }

void zkb200_set_profile(zkb200_ctx* ctx, int on) { for (auto& d : ctx->devs) d->profile = on != 0; }
unsigned long long zkb200_launch_count(void) { return g_kernel_launches.load(); }
int zkb200_last_stage_times(zkb200_ctx* ctx, const char** names, float* ms, int cap) {
  int n = (int)ctx->c.stage_ms.size();
  for (int i = 0; i < n && i < cap; i++) {
    if (names) names[i] = ctx->c.stage_ms[i].first.c_str();
    if (ms) ms[i] = ctx->c.stage_ms[i].second;
  }
  return n;
}

// ---- kernel-level entry points ----------------------------------------------------------------
int zkb200_coset_lde(zkb200_ctx* ctx, const uint32_t* in, uint32_t* out, unsigned log_n, size_t width, unsigned log_blowup,
                     uint32_t shift) {
  return guarded(ctx, [&] {
    std::lock_guard<std::mutex> lock(ctx->c.lanes[0].mu);
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    if (log_n + log_blowup > 24) throw std::runtime_error("zkb200: LDE height exceeds 2^24");
    size_t n = (size_t)1 << log_n;
    Lane& L = ctx->c.lanes[0];
    if (L.keep.size() > 32) { ZKB_CUDA(cudaStreamSynchronize(L.stream)); L.keep.clear(); }
    coset_lde_batch(ctx->c.tables, in, n, out, n << log_blowup, log_n, width, log_blowup, fp_from_canonical(shift), L.stream, L.keep);
  });
}
int zkb200_ntt(zkb200_ctx* ctx, const uint32_t* in, uint32_t* out, unsigned log_n, size_t width, int inverse, int bitrev_out) {
  return guarded(ctx, [&] {
    std::lock_guard<std::mutex> lock(ctx->c.lanes[0].mu);
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    if (log_n > 24) throw std::runtime_error("zkb200: NTT size exceeds 2^24");
    ntt_batch(ctx->c.tables, in, out, log_n, width, inverse != 0, bitrev_out != 0, ctx->c.lanes[0].stream);
  });
}
int zkb200_mmcs_root(zkb200_ctx* ctx, const uint32_t* const* mats, const unsigned* log_heights, const size_t* widths, int n,
                     uint32_t root_out[8]) {
  return guarded(ctx, [&] {
    std::lock_guard<std::mutex> lock(ctx->c.lanes[0].mu);
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    ctx->c.lanes[0].arena.reset();
    std::vector<MatRef> refs;
    for (int i = 0; i < n; i++) refs.push_back(MatRef{mats[i], (u32)widths[i], log_heights[i]});
    DigestLayers layers;
    merkle_build(refs, ctx->c.lanes[0].arena, layers, ctx->c.lanes[0].d_small, ctx->c.lanes[0].stream);
    ZKB_CUDA(cudaMemcpyAsync(ctx->c.lanes[0].h_small, ctx->c.lanes[0].d_small, 32, cudaMemcpyDeviceToHost, ctx->c.lanes[0].stream));
    ZKB_CUDA(cudaStreamSynchronize(ctx->c.lanes[0].stream));
    for (int i = 0; i < 8; i++) root_out[i] = fp_to_canonical(fp_raw(ctx->c.lanes[0].h_small[i]));
  });
}
int zkb200_poseidon2_permute_batch(zkb200_ctx* ctx, uint32_t* states, size_t n) {
  return guarded(ctx, [&] {
    std::lock_guard<std::mutex> lock(ctx->c.lanes[0].mu);
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    permute_batch(states, n, ctx->c.lanes[0].stream);
  });
}
int zkb200_permutation_trace(zkb200_ctx* ctx, const char* chip, const uint32_t* prep, const uint32_t* main_trace, size_t height,
                             const uint32_t alpha[4], const uint32_t beta[4], uint32_t* out, uint32_t local_sum_out[4]) {
  return guarded(ctx, [&] {
    std::lock_guard<std::mutex> lock(ctx->c.lanes[0].mu);
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    const ChipInfo* c = ctx->c.machine.find(chip);
    if (!c) throw std::runtime_error(std::string("zkb200: unknown chip ") + chip);
    permutation_trace(ctx->c.machine, *c, prep, main_trace, height, ef_from_canon(alpha), ef_from_canon(beta), out, ctx->c.lanes[0].d_small,
                      ctx->c.lanes[0].stream);
    ZKB_CUDA(cudaMemcpyAsync(ctx->c.lanes[0].h_small, ctx->c.lanes[0].d_small, 16, cudaMemcpyDeviceToHost, ctx->c.lanes[0].stream));
    ZKB_CUDA(cudaStreamSynchronize(ctx->c.lanes[0].stream));
    for (int i = 0; i < 4; i++) local_sum_out[i] = fp_to_canonical(fp_raw(ctx->c.lanes[0].h_small[i]));
  });
}
int zkb200_derive_multiplicities(zkb200_ctx* ctx, const char* receiver, const uint32_t* receiver_prep, size_t receiver_height,
                                 const zkb200_table* senders, int n_senders, uint32_t* out, uint64_t* n_lookups_out) {
  return guarded(ctx, [&] {
    std::lock_guard<std::mutex> lock(ctx->c.lanes[0].mu);
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    const ChipInfo* r = ctx->c.machine.find(receiver);
    if (!r) throw std::runtime_error(std::string("zkb200: unknown chip ") + receiver);
    if (!receiver_prep || !out) throw std::runtime_error("zkb200: derive_multiplicities: null receiver table");
    if (n_senders > 0 && !senders) throw std::runtime_error("zkb200: derive_multiplicities: null sender list");
    std::vector<DeriveSender> snd;
    for (int i = 0; i < n_senders; i++) {
      const ChipInfo* c = ctx->c.machine.find(senders[i].chip);
      if (!c) throw std::runtime_error(std::string("zkb200: unknown chip ") + senders[i].chip);
      if (senders[i].height && !senders[i].main_trace) throw std::runtime_error(std::string("zkb200: derive_multiplicities: no main trace for ") + senders[i].chip);
      if (c->prep_width && !senders[i].prep && senders[i].height) throw std::runtime_error(std::string("zkb200: derive_multiplicities: no preprocessed trace for ") + senders[i].chip);
      snd.push_back({c, DeriveTable{senders[i].prep, senders[i].main_trace, senders[i].height}});
    }
    const u64 n = derive_multiplicities(ctx->c.machine, *r, receiver_prep, receiver_height, snd, out, ctx->c.lanes[0].stream);
    if (n_lookups_out) *n_lookups_out = n;
  });
}
int zkb200_quotient(zkb200_ctx* ctx, const char* chip, unsigned log_n, const uint32_t* prep_lde, const uint32_t* main_lde,
                    const uint32_t* perm_lde, const uint32_t perm_alpha[4], const uint32_t perm_beta[4], const uint32_t local_sum[4],
                    const uint32_t global_sum[14], const uint32_t alpha[4], const uint32_t* pv, size_t npv, uint32_t* out) {
  return guarded(ctx, [&] {
    std::lock_guard<std::mutex> lock(ctx->c.lanes[0].mu);
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    ctx->c.lanes[0].arena.reset();
    const ChipInfo* c = ctx->c.machine.find(chip);
    if (!c) throw std::runtime_error(std::string("zkb200: unknown chip ") + chip);
    std::vector<u32> pvm(npv ? npv : 1, 0);
    for (size_t i = 0; i < npv; i++) pvm[i] = fp_from_canonical(pv[i]).v;
    QuotientInputs in;
    in.prep_lde = prep_lde; in.main_lde = main_lde; in.perm_lde = perm_lde;
    in.lde_h = (size_t)1 << (log_n + ctx->c.machine.log_blowup);
    in.log_n = log_n;
    in.perm_alpha = ef_from_canon(perm_alpha); in.perm_beta = ef_from_canon(perm_beta);
    in.local_sum = ef_from_canon(local_sum); in.alpha = ef_from_canon(alpha);
    for (int i = 0; i < 14; i++) in.global_sum[i] = fp_from_canonical(global_sum[i]).v;
    in.pub_dev = ctx->c.lanes[0].arena.push(pvm.data(), pvm.size());
    quotient_values(ctx->c.machine, *c, ctx->c.tables, in, out, ctx->c.lanes[0].stream);
    ZKB_CUDA(cudaStreamSynchronize(ctx->c.lanes[0].stream));
  });
}
int zkb200_fri_fold(zkb200_ctx* ctx, const uint32_t* in, size_t m, const uint32_t beta[4], const uint32_t* ro_next, uint32_t* out) {
  return guarded(ctx, [&] {
    std::lock_guard<std::mutex> lock(ctx->c.lanes[0].mu);
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    fri_fold(ctx->c.tables, in, m, ef_from_canon(beta), ro_next, out, ctx->c.lanes[0].stream);
  });
}
int zkb200_grind(zkb200_ctx* ctx, const uint32_t challenger[34], unsigned bits, uint32_t* witness_out) {
  return guarded(ctx, [&] {
    std::lock_guard<std::mutex> lock(ctx->c.lanes[0].mu);
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    Challenger ch;
    ch.load(challenger);
    u32 st[16];
    for (int i = 0; i < 16; i++) st[i] = ch.state[i].v;
    for (unsigned i = 0; i < ch.n_in; i++) st[i] = ch.in_buf[i].v;
    *witness_out = grind_witness(st, ch.n_in, bits, ctx->c.lanes[0].d_small + 8192, ctx->c.lanes[0].stream);
  });
}
int zkb200_transpose(zkb200_ctx* ctx, const uint32_t* in, uint32_t* out, size_t height, size_t width, int to_colmajor) {
  return guarded(ctx, [&] {
    std::lock_guard<std::mutex> lock(ctx->c.lanes[0].mu);
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    if (to_colmajor) transpose_to_colmajor(in, out, height, width, ctx->c.lanes[0].stream);
    else transpose_to_rowmajor(in, out, height, width, ctx->c.lanes[0].stream);
  });
}
int zkb200_convert(zkb200_ctx* ctx, uint32_t* data, size_t n, int to_montgomery) {
  return guarded(ctx, [&] {
    std::lock_guard<std::mutex> lock(ctx->c.lanes[0].mu);
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    if (to_montgomery) to_monty_inplace(data, n, ctx->c.lanes[0].stream);
    else from_monty_inplace(data, n, ctx->c.lanes[0].stream);
  });
}
int zkb200_alu_trace_width(const char* chip) {
  if (!strcmp(chip, "Global")) return GLOBAL_WIDTH;
  const int id = alu_chip_by_name(chip);
  return id < 0 ? -1 : alu_width(id);
}
int zkb200_generate_alu_trace(zkb200_ctx* ctx, const char* chip, const void* events, size_t n_events,
                              unsigned log_height, uint32_t* out, int col_major) {
  return guarded(ctx, [&] {
    static_assert(sizeof(zkb200_alu_event) == 28 && sizeof(zkb200_flow_event) == 28, "event records are seven 32-bit words");
    static_assert(sizeof(zkb200_cpu_event) == CPU_EVENT_WORDS * 4, "cpu event records are 28 32-bit words");
    const bool is_global = !strcmp(chip, "Global");          // the one table that is not row-local: lift, curve-point scan, finish
    const int id = is_global ? -1 : alu_chip_by_name(chip);
    if (id < 0 && !is_global) throw std::runtime_error(std::string("zkb200: generate_alu_trace: no row filler for chip ") + chip);
    if (log_height > 30) throw std::runtime_error("zkb200: generate_alu_trace: log_height out of range");
    const size_t height = (size_t)1 << log_height;
    if ((is_global ? n_events : ceil_div(n_events, (size_t)alu_events_per_row(id))) > height)
      throw std::runtime_error("zkb200: generate_alu_trace: more events than rows (fixed log2 rows is too small)");
    std::lock_guard<std::mutex> lock(ctx->c.lanes[0].mu);
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    cudaStream_t s = ctx->c.lanes[0].stream;
    const u32* ev = reinterpret_cast<const u32*>(events);
    DevBuf staged;
    cudaPointerAttributes attr;
    bool on_device = false;
    if (n_events && cudaPointerGetAttributes(&attr, events) == cudaSuccess)
      on_device = attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
    else cudaGetLastError();
    if (n_events && !on_device) {
      const size_t ew = is_global ? (size_t)GLOBAL_EVENT_WORDS : (size_t)alu_event_words(id);
      staged = DevBuf(n_events * ew, s);
      ZKB_CUDA(cudaMemcpyAsync(staged.p, events, n_events * ew * sizeof(u32), cudaMemcpyHostToDevice, s));
      ev = staged.p;
    }
    if (is_global) global_trace(ev, n_events, height, out, col_major != 0, s);
    else alu_trace(id, ev, n_events, height, out, col_major != 0, s);
    ZKB_CUDA(cudaStreamSynchronize(s));
  });
}
int zkb200_keccak_sponge_trace_width(void) { return KS_WIDTH; }
int zkb200_generate_keccak_sponge_trace(zkb200_ctx* ctx, const zkb200_keccak_block* blocks, size_t n_blocks,
                                        unsigned log_height, uint32_t* out, int col_major) {
  return guarded(ctx, [&] {
    static_assert(sizeof(zkb200_keccak_block) == KS_REC_WORDS * 4, "block records are 384 32-bit words");
    if (log_height > 30) throw std::runtime_error("zkb200: generate_keccak_sponge_trace: log_height out of range");
    const size_t height = (size_t)1 << log_height;
    if (n_blocks * KS_ROUNDS > height) throw std::runtime_error("zkb200: generate_keccak_sponge_trace: more rows than 2^log_height (fixed log2 rows is too small)");
    std::lock_guard<std::mutex> lock(ctx->c.lanes[0].mu);
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    cudaStream_t s = ctx->c.lanes[0].stream;
    const u32* ev = reinterpret_cast<const u32*>(blocks);
    DevBuf staged, tmp;
    cudaPointerAttributes attr;
    bool on_device = false;
    if (n_blocks && cudaPointerGetAttributes(&attr, blocks) == cudaSuccess)
      on_device = attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged;
    else cudaGetLastError();
    if (n_blocks && !on_device) {
      staged = DevBuf(n_blocks * KS_REC_WORDS, s);
      ZKB_CUDA(cudaMemcpyAsync(staged.p, blocks, n_blocks * sizeof(zkb200_keccak_block), cudaMemcpyHostToDevice, s));
      ev = staged.p;
    }
    if (col_major) keccak_sponge_trace(ev, n_blocks, height, out, s);
    else {
      // the kernel writes columns (coalesced); the RowMajorMatrix layout is one layout change away
      tmp = DevBuf(height * KS_WIDTH, s);
      keccak_sponge_trace(ev, n_blocks, height, tmp.p, s);
      transpose_to_rowmajor(tmp.p, out, height, KS_WIDTH, s);
    }
    ZKB_CUDA(cudaStreamSynchronize(s));
  });
}
// tools/h2d_probe.py: how fast do column slices of a pinned row-major matrix cross PCIe?
//   mode 0: 2-D DMA (cudaMemcpy2DAsync) of `seg_bytes`-wide slices, `nstreams` copies in flight
//   mode 1: the pull kernel (SM-initiated reads of mapped memory), `nstreams` = CTAs
//   mode 2: contiguous DMA of the whole buffer in `nstreams` concurrent parts
int zkb200_h2d_probe(zkb200_ctx* ctx, int mode, size_t row_bytes, size_t rows, size_t seg_bytes, int nstreams, float* ms_out) {
  return guarded(ctx, [&] {
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
    const size_t total = row_bytes * rows;
    static u32* host = nullptr; static size_t host_cap = 0;
    if (host_cap < total) { if (host) cudaFreeHost(host); ZKB_CUDA(cudaMallocHost((void**)&host, total)); host_cap = total; memset(host, 1, total); }
    u32* dev = nullptr;
    ZKB_CUDA(cudaMalloc((void**)&dev, total));
    std::vector<cudaStream_t> st(std::max(1, nstreams));
    if (mode != 1) for (auto& s : st) ZKB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    ZKB_CUDA(cudaEventCreate(&e0)); ZKB_CUDA(cudaEventCreate(&e1));
    ZKB_CUDA(cudaDeviceSynchronize());
    cudaStream_t main_s = ctx->c.lanes[0].stream;
    ZKB_CUDA(cudaEventRecord(e0, main_s));
    if (mode == 1) {
      PullPiece pp;
      pp.base = host; pp.word_off = 0; pp.pitch = row_bytes / 4; pp.dst = dev; pp.rows = rows; pp.cols = row_bytes / 4;
      pp.col_tiles = pull_piece_col_tiles(pp.cols); pp.tile_begin = 0;
      u32* work = nullptr;
      ZKB_CUDA(cudaMalloc((void**)&work, sizeof(PullPiece) + 256));
      ZKB_CUDA(cudaMemcpy(work, &pp, sizeof(PullPiece), cudaMemcpyHostToDevice));
      ZKB_CUDA(cudaMemset((char*)work + sizeof(PullPiece), 0, 4));
      ZKB_CUDA(cudaEventRecord(e0, main_s));
      pull_shard(reinterpret_cast<const PullPiece*>(work), 1, pull_piece_tiles(rows, pp.cols), (u32*)((char*)work + sizeof(PullPiece)), nstreams < 0 ? -nstreams : nstreams, nstreams < 0, main_s);
      ZKB_CUDA(cudaEventRecord(e1, main_s)); ZKB_CUDA(cudaEventSynchronize(e1));
      cudaFree(work);
    } else {
      std::vector<cudaEvent_t> done(st.size());
      for (size_t i = 0; i < st.size(); i++) { ZKB_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming)); ZKB_CUDA(cudaStreamWaitEvent(st[i], e0, 0)); }
      if (mode == 0) {
        size_t k = 0;
        for (size_t c0 = 0; c0 < row_bytes; c0 += seg_bytes, k++) {
          const size_t w = std::min(seg_bytes, row_bytes - c0);
          ZKB_CUDA(cudaMemcpy2DAsync((char*)dev + c0 * rows, w, (char*)host + c0, row_bytes, w, rows, cudaMemcpyHostToDevice, st[k % st.size()]));
        }
      } else {
        const size_t part = (total / st.size() + 255) & ~(size_t)255;
        for (size_t i = 0; i < st.size(); i++) {
          const size_t off = i * part, len = off < total ? std::min(part, total - off) : 0;
          if (len) ZKB_CUDA(cudaMemcpyAsync((char*)dev + off, (char*)host + off, len, cudaMemcpyHostToDevice, st[i]));
        }
      }
      for (size_t i = 0; i < st.size(); i++) { ZKB_CUDA(cudaEventRecord(done[i], st[i])); ZKB_CUDA(cudaStreamWaitEvent(main_s, done[i], 0)); }
      ZKB_CUDA(cudaEventRecord(e1, main_s));
      ZKB_CUDA(cudaEventSynchronize(e1));
      for (auto e : done) cudaEventDestroy(e);
    }
    ZKB_CUDA(cudaEventElapsedTime(ms_out, e0, e1));
    if (mode != 1) for (auto& s : st) cudaStreamDestroy(s);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dev);
  });
}
// CPU-side check of the run-time code generator: parse a ZKMD descriptor, generate every chip's constraint
// kernel and compile it with NVRTC for sm_100a (no GPU needed).  Returns the number of chips compiled, or -1
// (zkb200_last_error tells why); total cubin bytes in *bytes_out.
int zkb200_codegen_compile_check(const uint32_t* desc, size_t n_words, size_t* bytes_out) {
  int count = -1;
  guarded(nullptr, [&] {
    MachineInfo m;
    m.parse(desc, n_words);
    // chips compile in parallel (NVRTC is thread safe); results land in the kernel cache on disk
    std::vector<size_t> bytes(m.chips.size(), 0);
    std::vector<std::string> errs(m.chips.size());
    std::atomic<size_t> next{0};
    auto work = [&] {
      for (size_t i = next++; i < m.chips.size(); i = next++) {
        bytes[i] = quotient_codegen_compile_only(m.chips[i]);
        if (!bytes[i]) errs[i] = quotient_codegen_last_error();
      }
    };
    std::vector<std::thread> th;
    const unsigned nt = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)m.chips.size()));
    for (unsigned t = 0; t < nt; t++) th.emplace_back(work);
    for (auto& t : th) t.join();
    size_t total = 0;
    for (size_t i = 0; i < m.chips.size(); i++) {
      if (!bytes[i]) throw std::runtime_error("zkb200: codegen for chip " + m.chips[i].name + ": " + errs[i]);
      total += bytes[i];
    }
    if (bytes_out) *bytes_out = total;
    count = (int)m.chips.size();
  });
  return count;
}
// constraint-kernel launches so far: generated (NVRTC) and interpreter
void zkb200_quotient_launch_counts(unsigned long long* generated, unsigned long long* interpreted) {
  if (generated) *generated = g_quotient_generated_launches.load();
  if (interpreted) *interpreted = g_quotient_interpreter_launches.load();
}
int zkb200_set_option(const char* key, long value) {
  const std::string k(key ? key : "");
  if (k == "ntt_k2") { g_ntt_force_k2 = (int)value; return 0; }
  if (k == "eval_v2") { g_eval_v2 = (int)value; return 0; }
  if (k == "ntt_lean") { g_ntt_lean = (int)value; return 0; }
  if (k == "quotient_codegen") { g_quotient_codegen = (int)value; return 0; }
  if (k == "k4b_rows") { g_k4b_rows = (int)value; return 0; }
  if (k == "qk_block") { g_qk_block = (int)value; return 0; }
  if (k == "logup_codegen") { g_logup_codegen = (int)value; return 0; }
  g_err = "zkb200: unknown option " + k;
  return 1;
}
int zkb200_sync(zkb200_ctx* ctx) {
  return guarded(ctx, [&] {
    for (auto& d : ctx->devs) { ZKB_CUDA(cudaSetDevice(d->device)); ZKB_CUDA(cudaDeviceSynchronize()); }
    ZKB_CUDA(cudaSetDevice(ctx->c.device));
  });
}

}  // extern "C"

// K6: event -> row kernels for the core ALU and control-flow chips (seven-word event records).  One thread fills one row into a shared-memory
// tile (row stride odd, so that the column-wise read-back is conflict-free); the CTA then stores
// the tile with fully coalesced writes in either layout: its 128 rows are one contiguous block of
// the row-major matrix, and 128 consecutive elements of every column of the column-major one.
// HBM-bound by construction: 28 bytes read and 4 * width bytes written per row.
#include "tracegen.h"
#include <cstring>

namespace zkb {

__constant__ u32 d_inv255[256];

__constant__ GlobalConsts d_global_consts;

void tracegen_upload_constants() {
  u32 h[256];
  alu_build_inv255(h);
  ZKB_CUDA(cudaMemcpyToSymbol(d_inv255, h, sizeof(h)));
  static const GlobalConsts g = [] { GlobalConsts k; global_build_consts(k); return k; }();
  ZKB_CUDA(cudaMemcpyToSymbol(d_global_consts, &g, sizeof(g)));
}

constexpr int TG_ROWS = 128;   // threads per CTA of the KeccakSponge kernel

// the three phases are AluCta<CHIP> in tracegen.cuh (also walked on the host by tests/hostcheck)
template <int CHIP>
__global__ void __launch_bounds__(AluCta<CHIP>::R) alu_rows_kernel(const u32* __restrict__ events, size_t n, size_t height,
                                                                   u32* __restrict__ out, int col_major) {
  using C = AluCta<CHIP>;
  __shared__ u32 ev_s[C::R * C::RW];
  __shared__ u32 tile[C::R * C::WP];
  C::load(threadIdx.x, blockIdx.x, events, n, ev_s);
  __syncthreads();
  C::fill(threadIdx.x, blockIdx.x, n, ev_s, tile, d_inv255);
  __syncthreads();
  C::store(threadIdx.x, blockIdx.x, height, tile, out, col_major);
}

void alu_trace(int chip, const u32* events_dev, size_t n, size_t height, u32* out, bool col_major, cudaStream_t s) {
  if (!height) return;
  if (chip < 0 || chip >= ALU_NCHIPS) throw std::runtime_error("zkb200: alu_trace: unknown chip");
  if (ceil_div(n, (size_t)alu_events_per_row(chip)) > height) throw std::runtime_error("zkb200: alu_trace: more events than rows");
  const int cm = col_major ? 1 : 0;
#define ZKB_ALU_CASE(C) \
  case C: alu_rows_kernel<C><<<(unsigned)ceil_div(height, (size_t)AluCta<C>::R), AluCta<C>::R, 0, s>>>(events_dev, n, height, out, cm); break;
  switch (chip) {
    ZKB_ALU_CASE(ALU_ADDSUB) ZKB_ALU_CASE(ALU_BITWISE) ZKB_ALU_CASE(ALU_LT) ZKB_ALU_CASE(ALU_SLL) ZKB_ALU_CASE(ALU_SR)
    ZKB_ALU_CASE(ALU_CLOCLZ) ZKB_ALU_CASE(ALU_BRANCH) ZKB_ALU_CASE(ALU_JUMP) ZKB_ALU_CASE(ALU_MOVCOND) ZKB_ALU_CASE(ALU_MUL)
    ZKB_ALU_CASE(ALU_MEMINSTR) ZKB_ALU_CASE(ALU_MEMLOCAL) ZKB_ALU_CASE(ALU_CPU) ZKB_ALU_CASE(ALU_MISC) ZKB_ALU_CASE(ALU_DIVREM)
    ZKB_ALU_CASE(ALU_SYSCALL_CORE) ZKB_ALU_CASE(ALU_SYSCALL_PRECOMPILE) ZKB_ALU_CASE(ALU_SYSCALL_INSTRS)
    ZKB_ALU_CASE(ALU_MEMGLOBAL_INIT) ZKB_ALU_CASE(ALU_MEMGLOBAL_FINALIZE)
    default: throw std::runtime_error("zkb200: alu_trace: unknown chip");
  }
#undef ZKB_ALU_CASE
  ZKB_CHECK_LAUNCH();
}

// K6b: KeccakSponge rows, one thread per row (block = row / 24, round = row % 24), stored column-major: the 32
// rows of a warp are 32 consecutive words of every column.  HBM-write bound: 1536 bytes of record per 24 rows read
// (L2 hits for 23 of them), 4 * 3531 bytes per row written.
struct KsColStore {
  u32* p; u32 h;      // one 32 x 32 -> 64 multiply-add per address
  __device__ __forceinline__ void operator()(int col, u32 v) { p[(u64)((u32)col) * h] = v; }
};
__global__ void __launch_bounds__(TG_ROWS) keccak_sponge_rows_kernel(const u32* __restrict__ recs, size_t n_blocks, size_t height,
                                                                     u32* __restrict__ out) {
  const size_t row = (size_t)blockIdx.x * TG_ROWS + threadIdx.x;
  if (row >= height) return;
  const size_t b = row / KS_ROUNDS;
  KsColStore st{out + row, (u32)height};
  ks_fill_row(b < n_blocks ? recs + b * KS_REC_WORDS : nullptr, (u32)(row % KS_ROUNDS), st);
}

void keccak_sponge_trace(const u32* blocks_dev, size_t n_blocks, size_t height, u32* out_colmajor, cudaStream_t s) {
  if (!height) return;
  if (n_blocks * KS_ROUNDS > height) throw std::runtime_error("zkb200: keccak_sponge_trace: more rows than the table holds");
  keccak_sponge_rows_kernel<<<ceil_div(height, TG_ROWS), TG_ROWS, 0, s>>>(blocks_dev, n_blocks, height, out_colmajor);
  ZKB_CHECK_LAUNCH();
}

// K6c: the Global chip (tracegen_global.cuh): lift every event to its curve point, scan the points under the curve addition,
// write the accumulation columns.  One thread per event / chunk / row; compute-bound (a few thousand field products per event
// in the lift, about a thousand per curve addition), the rows are 99 words.
constexpr int TG_GLOBAL_THREADS = 128;
__global__ void __launch_bounds__(TG_GLOBAL_THREADS) global_lift_kernel(const u32* __restrict__ events, size_t n, GlobalOut out,
                                                                        u32* __restrict__ points) {
  const size_t row = (size_t)blockIdx.x * TG_GLOBAL_THREADS + threadIdx.x;
  if (row == 0) curve_store(points, d_global_consts.start);
  if (row < n) global_lift_row(events + GLOBAL_EVENT_WORDS * row, row, d_global_consts, out, points);
}
__global__ void __launch_bounds__(TG_GLOBAL_THREADS) global_total_kernel(const u32* __restrict__ pts, size_t n, size_t chunks,
                                                                         u32* __restrict__ totals) {
  const size_t t = (size_t)blockIdx.x * TG_GLOBAL_THREADS + threadIdx.x;
  if (t < chunks) global_chunk_total(pts, n, t, d_global_consts, totals);
}
__global__ void __launch_bounds__(TG_GLOBAL_THREADS) global_rescan_kernel(u32* __restrict__ pts, size_t n, size_t chunks,
                                                                          const u32* __restrict__ scanned_totals) {
  const size_t t = (size_t)blockIdx.x * TG_GLOBAL_THREADS + threadIdx.x;
  if (t < chunks) global_chunk_rescan(pts, n, t, d_global_consts, scanned_totals);
}
__global__ void __launch_bounds__(TG_GLOBAL_THREADS) global_finish_kernel(size_t n, size_t height, const u32* __restrict__ sums,
                                                                          GlobalOut out) {
  const size_t row = (size_t)blockIdx.x * TG_GLOBAL_THREADS + threadIdx.x;
  if (row < height) global_finish_row(row, n, sums, d_global_consts, out);
}
// inclusive scan of n points in place: chunk totals, the same scan over the totals, every chunk again from its prefix
static void global_scan(u32* pts, size_t n, cudaStream_t s) {
  if (n <= 1) return;
  const size_t chunks = ceil_div(n, (size_t)GLOBAL_SCAN_CHUNK);
  const unsigned grid = (unsigned)ceil_div(chunks, (size_t)TG_GLOBAL_THREADS);
  if (chunks == 1) {
    global_rescan_kernel<<<1, TG_GLOBAL_THREADS, 0, s>>>(pts, n, 1, nullptr);
    ZKB_CHECK_LAUNCH();
    return;
  }
  DevBuf totals(chunks * GLOBAL_POINT_WORDS, s);
  global_total_kernel<<<grid, TG_GLOBAL_THREADS, 0, s>>>(pts, n, chunks, totals.p);
  ZKB_CHECK_LAUNCH();
  global_scan(totals.p, chunks, s);
  global_rescan_kernel<<<grid, TG_GLOBAL_THREADS, 0, s>>>(pts, n, chunks, totals.p);
  ZKB_CHECK_LAUNCH();
}
void global_trace(const u32* events_dev, size_t n, size_t height, u32* out, bool col_major, cudaStream_t s) {
  if (!height) return;
  if (n > height) throw std::runtime_error("zkb200: global_trace: more events than rows");
  const GlobalOut o{out, col_major ? (size_t)1 : (size_t)GLOBAL_WIDTH, col_major ? height : (size_t)1};
  DevBuf points((n + 1) * GLOBAL_POINT_WORDS, s);
  global_lift_kernel<<<(unsigned)ceil_div(n ? n : (size_t)1, (size_t)TG_GLOBAL_THREADS), TG_GLOBAL_THREADS, 0, s>>>(events_dev, n, o, points.p);
  ZKB_CHECK_LAUNCH();
  global_scan(points.p, n + 1, s);
  global_finish_kernel<<<(unsigned)ceil_div(height, (size_t)TG_GLOBAL_THREADS), TG_GLOBAL_THREADS, 0, s>>>(n, height, points.p, o);
  ZKB_CHECK_LAUNCH();
}

int alu_chip_by_name(const char* name) {
  static const char* names[ALU_NCHIPS] = {"AddSub", "Bitwise", "Lt", "ShiftLeft", "ShiftRight", "CloClz", "Branch", "Jump", "MovCond", "Mul", "MemoryInstrs", "MemoryLocal", "Cpu", "MiscInstrs", "DivRem",
                                          "SyscallCore", "SyscallPrecompile", "SyscallInstrs", "MemoryGlobalInit", "MemoryGlobalFinalize"};
  for (int i = 0; i < ALU_NCHIPS; i++) if (!strcmp(name, names[i])) return i;
  return -1;
}

}  // namespace zkb

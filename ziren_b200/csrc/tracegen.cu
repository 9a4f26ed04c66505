// K6: event -> row kernels for the core ALU and control-flow chips (seven-word event records).  One thread fills one row into a shared-memory
// tile (row stride odd, so that the column-wise read-back is conflict-free); the CTA then stores
// the tile with fully coalesced writes in either layout: its 128 rows are one contiguous block of
// the row-major matrix, and 128 consecutive elements of every column of the column-major one.
// HBM-bound by construction: 28 bytes read and 4 * width bytes written per row.
#include "tracegen.h"
#include <cstring>

namespace zkb {

__constant__ u32 d_inv255[256];

void tracegen_upload_constants() {
  u32 h[256];
  alu_build_inv255(h);
  ZKB_CUDA(cudaMemcpyToSymbol(d_inv255, h, sizeof(h)));
}

constexpr int TG_ROWS = 128;   // rows per CTA = threads per CTA

template <int CHIP>
__global__ void __launch_bounds__(TG_ROWS) alu_rows_kernel(const u32* __restrict__ events, size_t n, size_t height,
                                                           u32* __restrict__ out, int col_major) {
  constexpr int W = alu_width(CHIP);
  constexpr int WP = W | 1;
  constexpr int EW = alu_event_words(CHIP);
  constexpr int EPR = alu_events_per_row(CHIP);
  constexpr int RW = EW * EPR;                    // event words per row
  static_assert((size_t)TG_ROWS * (RW + WP) * sizeof(u32) <= 48 * 1024, "events and row tile must fit the static shared-memory limit");
  __shared__ u32 ev_s[TG_ROWS * RW];
  __shared__ u32 tile[TG_ROWS * WP];
  const size_t row0 = (size_t)blockIdx.x * TG_ROWS;
  // the CTA's events are at most RW * 128 consecutive words: coalesced load, then one row's records per thread
  const size_t e0 = row0 * EPR;
  const size_t ev_words = e0 < n ? (n - e0 < (size_t)TG_ROWS * EPR ? (n - e0) * EW : (size_t)TG_ROWS * RW) : 0;
  for (u32 i = threadIdx.x; i < ev_words; i += TG_ROWS) ev_s[i] = events[e0 * EW + i];
  __syncthreads();
  u32* r = tile + threadIdx.x * WP;
  const size_t first = (row0 + threadIdx.x) * EPR;          // the row's first event
  if (first < n) fill_alu_row(CHIP, ev_s + RW * threadIdx.x, r, d_inv255, n - first < (size_t)EPR ? (int)(n - first) : EPR);
  else fill_alu_padding(CHIP, r);
  __syncthreads();
  const size_t rows = height - row0 < TG_ROWS ? height - row0 : TG_ROWS;
  if (col_major) {
    if (threadIdx.x < rows) {
#pragma unroll 4
      for (int c = 0; c < W; c++) out[(size_t)c * height + row0 + threadIdx.x] = tile[threadIdx.x * WP + c];
    }
  } else {
    u32* dst = out + row0 * W;
    for (u32 i = threadIdx.x; i < rows * W; i += TG_ROWS) dst[i] = tile[(i / W) * WP + (i % W)];
  }
}

void alu_trace(int chip, const u32* events_dev, size_t n, size_t height, u32* out, bool col_major, cudaStream_t s) {
  if (!height) return;
  if (ceil_div(n, (size_t)alu_events_per_row(chip)) > height) throw std::runtime_error("zkb200: alu_trace: more events than rows");
  const unsigned grid = ceil_div(height, TG_ROWS);
  const int cm = col_major ? 1 : 0;
  switch (chip) {
    case ALU_ADDSUB: alu_rows_kernel<ALU_ADDSUB><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    case ALU_BITWISE: alu_rows_kernel<ALU_BITWISE><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    case ALU_LT: alu_rows_kernel<ALU_LT><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    case ALU_SLL: alu_rows_kernel<ALU_SLL><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    case ALU_SR: alu_rows_kernel<ALU_SR><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    case ALU_CLOCLZ: alu_rows_kernel<ALU_CLOCLZ><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    case ALU_BRANCH: alu_rows_kernel<ALU_BRANCH><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    case ALU_JUMP: alu_rows_kernel<ALU_JUMP><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    case ALU_MOVCOND: alu_rows_kernel<ALU_MOVCOND><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    case ALU_MUL: alu_rows_kernel<ALU_MUL><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    case ALU_MEMINSTR: alu_rows_kernel<ALU_MEMINSTR><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    case ALU_MEMLOCAL: alu_rows_kernel<ALU_MEMLOCAL><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    case ALU_CPU: alu_rows_kernel<ALU_CPU><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    case ALU_MISC: alu_rows_kernel<ALU_MISC><<<grid, TG_ROWS, 0, s>>>(events_dev, n, height, out, cm); break;
    default: throw std::runtime_error("zkb200: alu_trace: unknown chip");
  }
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

int alu_chip_by_name(const char* name) {
  static const char* names[ALU_NCHIPS] = {"AddSub", "Bitwise", "Lt", "ShiftLeft", "ShiftRight", "CloClz", "Branch", "Jump", "MovCond", "Mul", "MemoryInstrs", "MemoryLocal", "Cpu", "MiscInstrs"};
  for (int i = 0; i < ALU_NCHIPS; i++) if (!strcmp(name, names[i])) return i;
  return -1;
}

}  // namespace zkb

#include "layout.h"
#include <algorithm>

namespace zkb {

// generic 32x32 tiled transpose of a (rows x cols) row-major array into (cols x rows).  Tiles are
// numbered along a 1-D grid (column tile fastest), so neither dimension meets the 65535-block limit
// of grid.y/z: the reference's default shards are 2^21 and 2^22 rows (crates/stark/src/opts.rs:42-50).
__global__ void transpose_kernel(const u32* __restrict__ in, size_t in_pitch, u32* __restrict__ out, size_t rows, size_t cols, unsigned col_tiles) {
  __shared__ u32 tile[32][33];
  const size_t c0 = (size_t)(blockIdx.x % col_tiles) * 32, r0 = (size_t)(blockIdx.x / col_tiles) * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    size_t r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = in[r * in_pitch + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    size_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][i];
  }
}
static void transpose(const u32* in, size_t in_pitch, u32* out, size_t rows, size_t cols, cudaStream_t s) {
  if (!rows || !cols) return;
  const size_t col_tiles = (cols + 31) / 32, row_tiles = (rows + 31) / 32;
  if (col_tiles * row_tiles > 0x7fffffffull) throw std::runtime_error("zkb200: matrix too large for one transpose launch");
  transpose_kernel<<<(unsigned)(col_tiles * row_tiles), dim3(32, 8), 0, s>>>(in, in_pitch, out, rows, cols, (unsigned)col_tiles);
  ZKB_CHECK_LAUNCH();
}
void transpose_to_colmajor(const u32* in, u32* out, size_t h, size_t w, cudaStream_t s) { transpose(in, w, out, h, w, s); }
void transpose_to_rowmajor(const u32* in, u32* out, size_t h, size_t w, cudaStream_t s) { transpose(in, h, out, w, h, s); }
void transpose_piece_to_colmajor(const u32* in, size_t in_pitch, u32* out, size_t h, size_t w, cudaStream_t s) { transpose(in, in_pitch, out, h, w, s); }

// ---- upload fused with the layout change ("pull") ------------------------------------------------
// The row-major host matrices lie in PINNED, device-mapped memory: ONE persistent kernel per shard reads
// them over PCIe itself, column piece by column piece, transposes tiles in shared memory, writes the
// column-major device traces and bumps a per-piece counter that the compute lane waits on
// (cuStreamWaitValue32).  No staging buffer, no second pass over HBM, and a column piece costs no more
// than a whole matrix.  Measured on a B200 (tools/h2d_probe.py): 51 GB/s against 55.6 GB/s for a
// contiguous DMA; a 2-D DMA of column slices is SM-driven as well and loses its SM slots to the compute
// lanes.  The kernel takes its few CTAs once (high-priority stream) and keeps them for the whole upload:
// 52 GB/s x ~2 us of latency is only ~100 KB in flight.
// Rows of odd width are not 128-byte aligned: every warp loads the ALIGNED 128-byte lines that cover its
// row segment (one extra line per 256 columns) instead of straddling two lines with every request.
constexpr int PULL_TR = 32, PULL_TC = 256, PULL_LINES = PULL_TC / 32 + 1;
// NW warps per CTA.  NW = 8: a small CTA that shares its SM with the compute kernels.  NW = 32 ("exclusive",
// ZKB200_PULL_EXCLUSIVE=1): 1024 threads and a 200 KB dynamic shared-memory request, so that the CTA owns
// its SM and its PCIe reads do not queue behind the compute kernels' global loads in that SM's load path.
template <int NW>
__global__ void __launch_bounds__(NW * 32) pull_shard_kernel(const PullPiece* __restrict__ pieces, int npieces,
                                                             unsigned long long total_tiles, u32* __restrict__ done) {
  constexpr int PULL_THREADS = NW * 32, RPW = PULL_TR / NW;       // rows per warp
  __shared__ u32 tile[PULL_TC][PULL_TR + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int pi = 0;
  for (unsigned long long tix = blockIdx.x; tix < total_tiles; tix += gridDim.x) {
    while (pi + 1 < npieces && tix >= pieces[pi + 1].tile_begin) pi++;
    const PullPiece p = pieces[pi];
    const unsigned long long local = tix - p.tile_begin;
    const size_t c0 = (size_t)(local % p.col_tiles) * PULL_TC, r0 = (size_t)(local / p.col_tiles) * PULL_TR;
    const size_t ncols = p.cols - c0 < (size_t)PULL_TC ? p.cols - c0 : (size_t)PULL_TC;
    // NW warps x RPW rows x 9 aligned lines: 9 * RPW independent loads per thread
    u32 v[RPW][PULL_LINES];
    long long first[RPW];
#pragma unroll
    for (int k = 0; k < RPW; k++) {
      const size_t r = r0 + warp * RPW + k;
      const size_t g0 = r * p.pitch + p.word_off + c0;       // word index of (r, c0) from the 128-byte aligned base
      const size_t a0 = g0 & ~(size_t)31;
      first[k] = (long long)a0 - (long long)g0;             // column (relative to c0) of the line's first word: -31 .. 0
#pragma unroll
      for (int j = 0; j < PULL_LINES; j++) {
        const long long cc = first[k] + 32 * j + lane;
        v[k][j] = (r < p.rows && cc >= 0 && cc < (long long)ncols) ? __ldcs(p.base + a0 + 32 * j + lane) : 0u;
      }
    }
#pragma unroll
    for (int k = 0; k < RPW; k++)
#pragma unroll
      for (int j = 0; j < PULL_LINES; j++) {
        const long long cc = first[k] + 32 * j + lane;
        if (cc >= 0 && cc < (long long)PULL_TC) tile[cc][warp * RPW + k] = v[k][j];
      }
    __syncthreads();
    // columns leave as 128-byte runs of 32 rows
    for (int cc = warp; cc < (int)ncols; cc += PULL_THREADS / 32) {
      const size_t r = r0 + lane;
      if (r < p.rows) p.dst[(c0 + cc) * p.rows + r] = tile[cc][lane];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(done + pi, 1u);
    }
  }
}
unsigned long long pull_piece_tiles(size_t rows, size_t cols) {
  return (unsigned long long)((cols + PULL_TC - 1) / PULL_TC) * ((rows + PULL_TR - 1) / PULL_TR);
}
unsigned pull_piece_col_tiles(size_t cols) { return (unsigned)((cols + PULL_TC - 1) / PULL_TC); }
void pull_set_device_attributes() {
  ZKB_CUDA(cudaFuncSetAttribute(pull_shard_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 190 * 1024));
}
void pull_shard(const PullPiece* pieces_dev, int npieces, unsigned long long total_tiles, u32* done, int ctas, bool exclusive,
                cudaStream_t s) {
  if (!npieces || !total_tiles) return;
  const unsigned grid = (unsigned)std::min<unsigned long long>(total_tiles, (unsigned long long)std::max(1, ctas));
  if (exclusive) pull_shard_kernel<32><<<grid, 1024, 190 * 1024, s>>>(pieces_dev, npieces, total_tiles, done);
  else pull_shard_kernel<8><<<grid, 256, 0, s>>>(pieces_dev, npieces, total_tiles, done);
  ZKB_CHECK_LAUNCH();
}

// see common.h: small parameter tables cross PCIe by SM loads from mapped pinned memory, not by the copy engine
__global__ void pull_words_kernel(u32* __restrict__ dst, const u32* __restrict__ src, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = __ldcs(src + i);
}
void pull_words(u32* dst_dev, const u32* src_pinned_mapped, size_t n_words, cudaStream_t s) {
  if (!n_words) return;
  const u32* mapped = nullptr;
  ZKB_CUDA(cudaHostGetDevicePointer((void**)&mapped, (void*)src_pinned_mapped, 0));
  const unsigned grid = (unsigned)std::min<size_t>(64, (n_words + 255) / 256);
  pull_words_kernel<<<grid, 256, 0, s>>>(dst_dev, mapped, n_words);
  ZKB_CHECK_LAUNCH();
}

__global__ void to_monty_kernel(u32* d, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) d[i] = fp_from_canonical(d[i]).v;
}
__global__ void from_monty_kernel(u32* d, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) d[i] = fp_to_canonical(fp_raw(d[i]));
}
void to_monty_inplace(u32* d, size_t n, cudaStream_t s) {
  if (!n) return;
  to_monty_kernel<<<ceil_div(n, 256), 256, 0, s>>>(d, n);
  ZKB_CHECK_LAUNCH();
}
void from_monty_inplace(u32* d, size_t n, cudaStream_t s) {
  if (!n) return;
  from_monty_kernel<<<ceil_div(n, 256), 256, 0, s>>>(d, n);
  ZKB_CHECK_LAUNCH();
}

}  // namespace zkb

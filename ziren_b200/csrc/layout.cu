#include "layout.h"

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

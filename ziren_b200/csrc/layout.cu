#include "layout.h"

namespace zkb {

// generic 32x32 tiled transpose of a (rows x cols) row-major array into (cols x rows)
__global__ void transpose_kernel(const u32* __restrict__ in, u32* __restrict__ out, size_t rows, size_t cols) {
  __shared__ u32 tile[32][33];
  size_t c0 = (size_t)blockIdx.x * 32, r0 = (size_t)blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    size_t r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = in[r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    size_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][i];
  }
}
static void transpose(const u32* in, u32* out, size_t rows, size_t cols, cudaStream_t s) {
  if (!rows || !cols) return;
  // grid.y is limited to 65535 blocks: walk the row dimension in slabs
  const size_t slab = (size_t)65535 * 32;
  for (size_t r0 = 0; r0 < rows; r0 += slab) {
    size_t nr = rows - r0 < slab ? rows - r0 : slab;
    dim3 grid(ceil_div(cols, 32), ceil_div(nr, 32));
    // the slab is a sub-range of rows: input offset r0*cols, output offset r0 within each column
    // (output leading dimension stays `rows`), so pass full `rows` through a shifted pointer
    if (r0 == 0 && nr == rows) {
      transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(in, out, rows, cols);
    } else {
      throw std::runtime_error("zkb200: matrix too tall for transpose");
    }
    ZKB_CHECK_LAUNCH();
  }
}
void transpose_to_colmajor(const u32* in, u32* out, size_t h, size_t w, cudaStream_t s) { transpose(in, out, h, w, s); }
void transpose_to_rowmajor(const u32* in, u32* out, size_t h, size_t w, cudaStream_t s) { transpose(in, out, w, h, s); }

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

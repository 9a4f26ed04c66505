// Layout changes between the reference's row-major host matrices and the column-major device
// layout (see DESIGN.md "Data layout in HBM").
#pragma once
#include "common.h"

namespace zkb {
// in: row-major h x w (device), out: column-major (w columns of h)
void transpose_to_colmajor(const u32* in, u32* out, size_t h, size_t w, cudaStream_t s);
// in: column-major, out: row-major h x w
// w columns of a row-major matrix whose rows are in_pitch words apart -> column-major h x w (column stride h)
void transpose_piece_to_colmajor(const u32* in, size_t in_pitch, u32* out, size_t h, size_t w, cudaStream_t s);
// Upload fused with the layout change (layout.cu): one persistent kernel pulls every column piece of a
// shard's pinned, device-mapped host traces over PCIe; done[i] counts the finished tiles of piece i.
struct PullPiece {
  const u32* base;                 // 128-byte aligned device-visible address at or below the matrix
  size_t word_off;                 // word offset of element (0, first column of the piece) from `base`
  size_t pitch;                    // words between rows
  u32* dst;                        // column-major piece: rows x cols, column stride `rows`
  size_t rows, cols;
  unsigned col_tiles;
  unsigned long long tile_begin;   // prefix sum of the tile counts
};
unsigned long long pull_piece_tiles(size_t rows, size_t cols);
unsigned pull_piece_col_tiles(size_t cols);
void pull_set_device_attributes();     // once per device
void pull_shard(const PullPiece* pieces_dev, int npieces, unsigned long long total_tiles, u32* done, int ctas, bool exclusive,
                cudaStream_t s);
void transpose_to_rowmajor(const u32* in, u32* out, size_t h, size_t w, cudaStream_t s);
// converts between canonical and Montgomery form in place
void to_monty_inplace(u32* d, size_t n, cudaStream_t s);
void from_monty_inplace(u32* d, size_t n, cudaStream_t s);
}  // namespace zkb

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
void transpose_to_rowmajor(const u32* in, u32* out, size_t h, size_t w, cudaStream_t s);
// converts between canonical and Montgomery form in place
void to_monty_inplace(u32* d, size_t n, cudaStream_t s);
void from_monty_inplace(u32* d, size_t n, cudaStream_t s);
}  // namespace zkb

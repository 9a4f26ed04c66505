// Trace generation on the GPU for the core ALU chips.
#pragma once
#include "common.h"
#include "tracegen.cuh"
#include "tracegen_keccak.cuh"
#include "tracegen_global.cuh"

namespace zkb {

void tracegen_upload_constants();   // once per device: the 1/d table of the Lt chip

// events: n records of 28 bytes (#[repr(C)] AluEvent) in DEVICE memory; out: height x width words,
// row-major (the RowMajorMatrix layout zkb200_commit takes) or column-major (the layout every kernel
// of this library works in).  Rows >= n are the chip's padding rows.
void alu_trace(int chip, const u32* events_dev, size_t n, size_t height, u32* out, bool col_major, cudaStream_t s);

// blocks: n_blocks records of KS_REC_WORDS words in DEVICE memory (include/zkb200.h, zkb200_keccak_block); out: height x
// 3531 words COLUMN-MAJOR; rows >= 24 * n_blocks are the chip's padding rows.
void keccak_sponge_trace(const u32* blocks_dev, size_t n_blocks, size_t height, u32* out_colmajor, cudaStream_t s);

// events: n GlobalLookupEvent records of 8 words in DEVICE memory; out: height x 99 words, row-major or column-major; rows >= n
// are the chip's dummy rows.  Scratch (the points and their running sums) is taken from the stream's memory pool.
void global_trace(const u32* events_dev, size_t n, size_t height, u32* out, bool col_major, cudaStream_t s);

int alu_chip_by_name(const char* name);   // MachineAir::name -> AluChip, -1 if not an ALU chip handled here

}  // namespace zkb

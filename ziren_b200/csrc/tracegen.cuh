// Row fillers of the core ALU chips (trace generation on the GPU).
// One AluEvent (crates/core/executor/src/events/instr.rs:11-26) becomes one row of the chip's
// RowMajorMatrix; the column order is the #[repr(C)] order of the chip's column struct.  What each
// column holds follows the chips' event_to_row (Rust, cited per chip) and its C++ twin
// crates/core/machine/include/*.hpp that the reference's `sys` feature calls; the formulation
// here is branch-free word arithmetic (carries, shifts and byte compares on whole 32/64-bit
// words) instead of the reference's per-byte loops.  Host and device run the same code
// (tests/hostcheck compiles it for the CPU and compares with the reference's C++).
#pragma once
#include "kb31.cuh"

namespace zkb {

enum AluChip : int { ALU_ADDSUB = 0, ALU_BITWISE = 1, ALU_LT = 2, ALU_SLL = 3, ALU_SR = 4, ALU_CLOCLZ = 5, ALU_BRANCH = 6, ALU_JUMP = 7,
                     ALU_MOVCOND = 8, ALU_MUL = 9, ALU_MEMINSTR = 10, ALU_MEMLOCAL = 11, ALU_CPU = 12, ALU_MISC = 13, ALU_DIVREM = 14, ALU_SYSCALL_CORE = 15,
                     ALU_SYSCALL_PRECOMPILE = 16, ALU_SYSCALL_INSTRS = 17, ALU_MEMGLOBAL_INIT = 18, ALU_MEMGLOBAL_FINALIZE = 19, ALU_NCHIPS = 20 };
// opcodes as numbered by crates/core/executor/src/opcode.rs:25-49 (#[repr(u8)])
enum : u32 { OP_ADD = 0, OP_SUB = 1, OP_SLL = 9, OP_SRL = 10, OP_SRA = 11, OP_ROR = 12, OP_SLT = 13, OP_SLTU = 14,
             OP_AND = 15, OP_OR = 16, OP_XOR = 17, OP_NOR = 18, OP_CLZ = 19, OP_CLO = 20, OP_BEQ = 21, OP_BGEZ = 22, OP_BGTZ = 23,
             OP_BLEZ = 24, OP_BLTZ = 25, OP_BNE = 26, OP_JUMP = 27, OP_JUMPI = 28, OP_JUMPDIRECT = 29, OP_MEQ = 50, OP_MNE = 51, OP_WSBH = 52 };

KB_HD constexpr int alu_width(int chip) {
  return chip == ALU_ADDSUB ? 19 : chip == ALU_BITWISE ? 18 : chip == ALU_LT ? 32 : chip == ALU_SLL ? 44 : chip == ALU_SR ? 67
       : chip == ALU_CLOCLZ ? 17 : chip == ALU_BRANCH ? 62 : chip == ALU_JUMP ? 66 : chip == ALU_MUL ? 58 : chip == ALU_MEMINSTR ? 79 : chip == ALU_MEMLOCAL ? 56 : chip == ALU_CPU ? 67 : chip == ALU_MISC ? 72
       : chip == ALU_DIVREM ? 106 : chip == ALU_SYSCALL_CORE || chip == ALU_SYSCALL_PRECOMPILE ? 11 : chip == ALU_SYSCALL_INSTRS ? 77
       : chip == ALU_MEMGLOBAL_INIT || chip == ALU_MEMGLOBAL_FINALIZE ? 111 : 32;
}
// 32-bit words per event record: seven for AluEvent / BranchEvent / JumpEvent / MovCondEvent, sixteen for CompAluEvent (Mul)
// and MemInstrEvent (MemoryInstrs), twenty-eight for the flattened CpuEvent + Instruction (zkb200_cpu_event)
// fifteen for MiscEvent (MiscInstrs), sixteen for CompAluEvent (DivRem), fourteen for SyscallEvent (SyscallCore, SyscallPrecompile,
// SyscallInstrs), six for the flattened MemoryInitializeFinalizeEvent (zkb200_memory_global_event)
KB_HD constexpr int alu_event_words(int chip) {
  return chip == ALU_CPU ? 28 : chip == ALU_MUL || chip == ALU_MEMINSTR || chip == ALU_DIVREM ? 16 : chip == ALU_MISC ? 15
       : chip == ALU_SYSCALL_CORE || chip == ALU_SYSCALL_PRECOMPILE || chip == ALU_SYSCALL_INSTRS ? 14
       : chip == ALU_MEMGLOBAL_INIT || chip == ALU_MEMGLOBAL_FINALIZE ? 6 : 7;
}
// events per row: MemoryLocal packs four seven-word MemoryLocalEvents into a row, every other chip has one event per row
KB_HD constexpr int alu_events_per_row(int chip) { return chip == ALU_MEMLOCAL ? 4 : 1; }

// AluEvent as laid out by #[repr(C)]: seven 32-bit words, the opcode in the low byte of word 2
struct AluEv { u32 pc, next_pc, opcode, hi, a, b, c; };
KB_HD AluEv alu_event_from_words(const u32* w) {
  AluEv e;
  e.pc = w[0]; e.next_pc = w[1]; e.opcode = w[2] & 0xffu; e.hi = w[3]; e.a = w[4]; e.b = w[5]; e.c = w[6];
  return e;
}

KB_HD u32 tg_top_bit(u32 x) {                                 // index of the highest set bit, x != 0
#if defined(__CUDA_ARCH__)
  return 31u - (u32)__clz((int)x);
#else
  return 31u - (u32)__builtin_clz(x);
#endif
}
KB_HD u32 tg_f(u32 x) { return fp_from_canonical(x).v; }      // F::from_canonical_u32 (x < p)
KB_HD u32 tg_b(bool x) { return x ? KB_ONE : 0u; }            // F::from_bool
KB_HD void tg_word(u32* r, u32 v) {                           // Word::from(u32): little-endian bytes
  r[0] = tg_f(v & 0xffu); r[1] = tg_f((v >> 8) & 0xffu); r[2] = tg_f((v >> 16) & 0xffu); r[3] = tg_f(v >> 24);
}
KB_HD void tg_long(u32* r, u64 v) { tg_word(r, (u32)v); tg_word(r + 4, (u32)(v >> 32)); }
KB_HD void tg_onehot(u32* r, int n, u32 hot) { for (int i = 0; i < n; i++) r[i] = tg_b((u32)i == hot); }
KB_HD void tg_bits(u32* r, int n, u32 v) { for (int i = 0; i < n; i++) r[i] = tg_b((v >> i) & 1u); }

// AddSubChip::event_to_row, crates/core/machine/src/alu/add_sub/mod.rs:152-171; AddOperation::populate,
// operations/add.rs:23-47.  Columns: pc, next_pc, add_operation{value[4], carry[3]}, operand_1[4],
// operand_2[4], is_add, is_sub.
KB_HD void fill_add_sub(const AluEv& e, u32* r) {
  const bool is_add = e.opcode == OP_ADD;
  const u32 x = is_add ? e.b : e.a, y = e.c;
  r[0] = tg_f(e.pc); r[1] = tg_f(e.next_pc);
  tg_word(r + 2, x + y);
  // carry out of the low 1, 2, 3 bytes
  r[6] = tg_b(((x & 0xffu) + (y & 0xffu)) >> 8);
  r[7] = tg_b(((x & 0xffffu) + (y & 0xffffu)) >> 16);
  r[8] = tg_b(((x & 0xffffffu) + (y & 0xffffffu)) >> 24);
  tg_word(r + 9, x);
  tg_word(r + 13, y);
  r[17] = tg_b(is_add);
  r[18] = tg_b(e.opcode == OP_SUB);
}

// BitwiseChip::event_to_row, crates/core/machine/src/alu/bitwise/mod.rs.  Columns: pc, next_pc, a[4],
// b[4], c[4], is_nor, is_xor, is_or, is_and.
KB_HD void fill_bitwise(const AluEv& e, u32* r) {
  r[0] = tg_f(e.pc); r[1] = tg_f(e.next_pc);
  tg_word(r + 2, e.a); tg_word(r + 6, e.b); tg_word(r + 10, e.c);
  r[14] = tg_b(e.opcode == OP_NOR); r[15] = tg_b(e.opcode == OP_XOR);
  r[16] = tg_b(e.opcode == OP_OR); r[17] = tg_b(e.opcode == OP_AND);
}

// LtChip::event_to_row, crates/core/machine/src/alu/lt/mod.rs.  Columns: pc, next_pc, is_slt, is_sltu,
// a[4], b[4], c[4], byte_flags[4], b_masked, c_masked, not_eq_inv, msb_b, msb_c, bit_b, bit_c, sltu,
// is_comp_eq, is_sign_eq, comparison_bytes[2].  inv255[d] = 1/d (Montgomery) for d = 1..255.
KB_HD void fill_lt(const AluEv& e, u32* r, const u32* inv255) {
  const bool slt = e.opcode == OP_SLT;
  r[0] = tg_f(e.pc); r[1] = tg_f(e.next_pc);
  r[2] = tg_b(slt); r[3] = tg_b(e.opcode == OP_SLTU);
  tg_word(r + 4, e.a); tg_word(r + 8, e.b); tg_word(r + 12, e.c);
  // SLT compares with the sign bits cleared (they are handled by bit_b / bit_c)
  const u32 bc = slt ? e.b & 0x7fffffffu : e.b, cc = slt ? e.c & 0x7fffffffu : e.c;
  const u32 diff = bc ^ cc;
  u32 hot = 4, bb = 0, cb = 0, inv = 0;
  if (diff) {
    // most significant differing byte
    hot = tg_top_bit(diff) >> 3;
    bb = (bc >> (8 * hot)) & 0xffu;
    cb = (cc >> (8 * hot)) & 0xffu;
    inv = bb > cb ? inv255[bb - cb] : KB_P - inv255[cb - bb];
  }
  tg_onehot(r + 16, 4, hot);
  r[20] = tg_f((e.b >> 24) & 0x7fu); r[21] = tg_f((e.c >> 24) & 0x7fu);
  r[22] = inv;
  const u32 msb_b = e.b >> 31, msb_c = e.c >> 31;
  r[23] = tg_b(msb_b); r[24] = tg_b(msb_c);
  r[25] = tg_b(msb_b && slt); r[26] = tg_b(msb_c && slt);
  r[27] = tg_b(bc < cc);
  r[28] = tg_b(diff == 0);
  r[29] = tg_b(!slt || msb_b == msb_c);
  r[30] = tg_f(bb); r[31] = tg_f(cb);
}

// ShiftLeft::event_to_row, crates/core/machine/src/alu/sll/mod.rs.  Columns: pc, next_pc, a[4], b[4],
// c[4], c_least_sig_byte[8], shift_by_n_bits[8], bit_shift_multiplier, bit_shift_result[4],
// bit_shift_result_carry[4], shift_by_n_bytes[4], is_real.
KB_HD void fill_shift_left(const AluEv& e, u32* r) {
  r[0] = tg_f(e.pc); r[1] = tg_f(e.next_pc);
  tg_word(r + 2, e.a); tg_word(r + 6, e.b); tg_word(r + 10, e.c);
  tg_bits(r + 14, 8, e.c);
  const u32 nbits = e.c & 7u, nbytes = (e.c & 31u) >> 3;
  tg_onehot(r + 22, 8, nbits);
  r[30] = tg_f(1u << nbits);
  // b * 2^nbits byte by byte: result bytes are those of the 64-bit shift, the carry after byte i
  // is what the low i+1 bytes push beyond them
  const u64 sh = (u64)e.b << nbits;
  tg_word(r + 31, (u32)sh);
  for (int i = 0; i < 4; i++) {
    const u64 low = (u64)(i == 3 ? e.b : (e.b & ((1u << (8 * (i + 1))) - 1u))) << nbits;
    r[35 + i] = tg_f((u32)(low >> (8 * (i + 1))));
  }
  tg_onehot(r + 39, 4, nbytes);
  r[43] = KB_ONE;
}

// ShiftRightChip::event_to_row, crates/core/machine/src/alu/sr/mod.rs.  Columns: pc, next_pc, b[4], c[4],
// shift_by_n_bits[8], shift_by_n_bytes[4], byte_shift_result[8], bit_shift_result[8],
// shr_carry_output_carry[8], shr_carry_output_shifted_byte[8], b_msb, c_least_sig_byte[8], is_srl,
// is_ror, is_sra, is_real.
KB_HD void fill_shift_right(const AluEv& e, u32* r) {
  r[0] = tg_f(e.pc); r[1] = tg_f(e.next_pc);
  tg_word(r + 2, e.b); tg_word(r + 6, e.c);
  const u32 nbits = e.c & 7u, nbytes = (e.c & 31u) >> 3;
  tg_onehot(r + 10, 8, nbits);
  tg_onehot(r + 18, 4, nbytes);
  // the operand widened to 64 bits: sign-extended (SRA), doubled (ROR) or zero-extended
  u64 wide = e.b;
  if (e.opcode == OP_SRA) wide = (u64)(int64_t)(int32_t)e.b;
  else if (e.opcode == OP_ROR) wide |= (u64)e.b << 32;
  const u64 by_bytes = wide >> (8 * nbytes);
  tg_long(r + 22, by_bytes);
  tg_long(r + 30, by_bytes >> nbits);
  // per byte: the bits shifted out (carry) and what stays (shifted byte)
  const u64 ones = 0x0101010101010101ull;
  const u64 carry_mask = ones * ((1u << nbits) - 1u);
  tg_long(r + 38, by_bytes & carry_mask);
  tg_long(r + 46, (by_bytes >> nbits) & (ones * (0xffu >> nbits)));
  r[54] = tg_b(e.b >> 31);
  tg_bits(r + 55, 8, e.c);
  r[63] = tg_b(e.opcode == OP_SRL); r[64] = tg_b(e.opcode == OP_ROR); r[65] = tg_b(e.opcode == OP_SRA);
  r[66] = KB_ONE;
}

// CloClzChip::generate_trace, crates/core/machine/src/alu/clo_clz/mod.rs:105-122.  Columns: pc, next_pc,
// a[4], b[4], bb[4], is_bb_zero, is_clz, is_real.
KB_HD void fill_clo_clz(const AluEv& e, u32* r) {
  const bool clz = e.opcode == OP_CLZ;
  const u32 bb = clz ? e.b : ~e.b;
  r[0] = tg_f(e.pc); r[1] = tg_f(e.next_pc);
  tg_word(r + 2, e.a); tg_word(r + 6, e.b); tg_word(r + 10, bb);
  r[14] = tg_b(bb == 0); r[15] = tg_b(clz); r[16] = KB_ONE;
}

// BranchEvent / JumpEvent (crates/core/executor/src/events/instr.rs:160-217, #[repr(C)]): seven words
// {pc, next_pc, next_next_pc, opcode, a, b, c}
struct FlowEv { u32 pc, next_pc, next_next_pc, opcode, a, b, c; };
KB_HD FlowEv flow_event_from_words(const u32* w) {
  FlowEv e;
  e.pc = w[0]; e.next_pc = w[1]; e.next_next_pc = w[2]; e.opcode = w[3] & 0xffu; e.a = w[4]; e.b = w[5]; e.c = w[6];
  return e;
}
// KoalaBearWordRangeChecker::populate, crates/core/machine/src/operations/koala_bear_word.rs:35-49: the bits
// of the most significant byte and the running conjunction of its low 2..7 bits (14 columns)
KB_HD void tg_range_checker(u32* r, u32 v) {
  const u32 top = v >> 24;
  tg_bits(r, 8, top);
  for (u32 j = 2; j <= 7; j++) r[6 + j] = tg_b((top & ((1u << j) - 1u)) == (1u << j) - 1u);
}
// BranchChip::event_to_row, crates/core/machine/src/control_flow/branch/trace.rs:94-137.  Columns
// (columns.rs): pc, next_pc[4], next_pc_range_checker[14], target_pc[4], next_next_pc[4],
// next_next_pc_range_checker[14], op_a_value[4], op_b_value[4], op_c_value[4], is_beq, is_bne, is_bltz,
// is_blez, is_bgtz, is_bgez, is_branching, a_gt_b, a_lt_b.
KB_HD void fill_branch(const FlowEv& e, u32* r) {
  r[0] = tg_f(e.pc);
  tg_word(r + 1, e.next_pc);
  tg_range_checker(r + 5, e.next_pc);
  tg_word(r + 19, e.next_pc + e.c);
  tg_word(r + 23, e.next_next_pc);
  tg_range_checker(r + 27, e.next_next_pc);
  tg_word(r + 41, e.a); tg_word(r + 45, e.b); tg_word(r + 49, e.c);
  const u32 op = e.opcode;
  r[53] = tg_b(op == OP_BEQ); r[54] = tg_b(op == OP_BNE); r[55] = tg_b(op == OP_BLTZ);
  r[56] = tg_b(op == OP_BLEZ); r[57] = tg_b(op == OP_BGTZ); r[58] = tg_b(op == OP_BGEZ);
  const bool eq = e.a == e.b, lt = (int32_t)e.a < (int32_t)e.b, gt = (int32_t)e.a > (int32_t)e.b;
  // taken when the relation the opcode names holds between a and b
  const bool want_eq = op == OP_BEQ || op == OP_BLEZ || op == OP_BGEZ;
  const bool want_lt = op == OP_BLTZ || op == OP_BLEZ || op == OP_BNE;
  const bool want_gt = op == OP_BGTZ || op == OP_BGEZ || op == OP_BNE;
  r[59] = tg_b((want_eq && eq) || (want_lt && lt) || (want_gt && gt));
  r[60] = tg_b(gt); r[61] = tg_b(lt);
}
// JumpChip::event_to_row, crates/core/machine/src/control_flow/jump/trace.rs:94-113.  Columns: pc, next_pc[4],
// next_pc_range_checker[14], next_next_pc[4], next_next_pc_range_checker[14], op_a_value[4], op_b_value[4],
// op_c_value[4], is_jump, is_jumpi, is_jumpdirect, op_a_range_checker[14].
KB_HD void fill_jump(const FlowEv& e, u32* r) {
  r[0] = tg_f(e.pc);
  tg_word(r + 1, e.next_pc);
  tg_range_checker(r + 5, e.next_pc);
  tg_word(r + 19, e.next_next_pc);
  tg_range_checker(r + 23, e.next_next_pc);
  tg_word(r + 37, e.a); tg_word(r + 41, e.b); tg_word(r + 45, e.c);
  r[49] = tg_b(e.opcode == OP_JUMP); r[50] = tg_b(e.opcode == OP_JUMPI); r[51] = tg_b(e.opcode == OP_JUMPDIRECT);
  tg_range_checker(r + 52, e.a);
}

// MovCondChip::event_to_row, crates/core/machine/src/misc/mov_cond/mod.rs:140-165; MovCondEvent
// (crates/core/executor/src/events/instr.rs:287-302) is {pc, next_pc, opcode, a, b, c, prev_a}.  Columns: pc,
// next_pc, op_a_value[4], prev_a_value[4], op_b_value[4], op_c_value[4], c_eq_0 = IsZeroWordOperation
// {is_zero_byte[4]{inverse, result}, is_lower_half_zero, is_upper_half_zero, result}
// (operations/is_zero_word.rs, is_zero.rs), is_mne, is_meq, is_wsbh.
KB_HD void fill_mov_cond(const u32* w, u32* r, const u32* inv255) {
  const u32 pc = w[0], next_pc = w[1], op = w[2] & 0xffu, a = w[3], b = w[4], c = w[5], prev_a = w[6];
  r[0] = tg_f(pc); r[1] = tg_f(next_pc);
  tg_word(r + 2, a); tg_word(r + 6, prev_a); tg_word(r + 10, b); tg_word(r + 14, c);
  for (int i = 0; i < 4; i++) {
    const u32 byte = (c >> (8 * i)) & 0xffu;
    r[18 + 2 * i] = inv255[byte];            // entry 0 of the table is 0
    r[19 + 2 * i] = tg_b(byte == 0);
  }
  r[26] = tg_b((c & 0xffffu) == 0); r[27] = tg_b((c >> 16) == 0); r[28] = tg_b(c == 0);
  r[29] = tg_b(op == OP_MNE); r[30] = tg_b(op == OP_MEQ); r[31] = tg_b(op == OP_WSBH);
}

// MulChip::event_to_row, crates/core/machine/src/alu/mul/mod.rs:235-336 (C++ twin include/mul.hpp).  Event: CompAluEvent
// (crates/core/executor/src/events/instr.rs:47-73) as 16 words {shard, clk, pc, next_pc, opcode, hi, a, b, c, hi_record{value,
// shard, timestamp, prev_value, prev_shard, prev_timestamp}, hi_record_is_real}.  Columns (58): pc, next_pc, hi[4], a[4], b[4],
// c[4], carry[8], product[8], b_msb, c_msb, b_sign_extend, c_sign_extend, is_mul, is_mult, is_multu, is_real, op_hi_access
// {prev_value[4], value[4], prev_shard, prev_clk, compare_clk, diff_16bit_limb, diff_8bit_limb}, hi_record_is_real, shard, clk.
// The operands are widened to 64 bits (sign-extended for MULT); byte k of the schoolbook product and the carry out of it
// come from the column sums of the 8 x 8 byte products.
constexpr int MUL_WIDTH = 58, COMP_EVENT_WORDS = 16;
enum : u32 { OP_MUL = 2, OP_MULT = 3, OP_MULTU = 4 };
KB_HD void fill_mul(const u32* e, u32* r) {
  const u32 shard = e[0], clk = e[1], pc = e[2], next_pc = e[3], op = e[4] & 0xffu, hi = e[5], a = e[6], b = e[7], c = e[8];
  const bool hi_real = (e[15] & 0xffu) != 0;          // a Rust bool: one byte, the other three are padding
  r[0] = tg_f(pc); r[1] = tg_f(next_pc);
  tg_word(r + 2, hi); tg_word(r + 6, a); tg_word(r + 10, b); tg_word(r + 14, c);
  const bool b_ext = op == OP_MULT && (b >> 31), c_ext = op == OP_MULT && (c >> 31);
  const u64 bw = b_ext ? (u64)(int64_t)(int32_t)b : (u64)b, cw = c_ext ? (u64)(int64_t)(int32_t)c : (u64)c;
  u32 carry = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    u32 sum = carry;
#pragma unroll
    for (int i = 0; i <= k; i++) sum += (u32)((bw >> (8 * i)) & 0xffu) * (u32)((cw >> (8 * (k - i))) & 0xffu);
    carry = sum >> 8;
    r[18 + k] = tg_f(carry);
    r[26 + k] = tg_f(sum & 0xffu);
  }
  r[34] = tg_b(b >> 31); r[35] = tg_b(c >> 31); r[36] = tg_b(b_ext); r[37] = tg_b(c_ext);
  r[38] = tg_b(op == OP_MUL); r[39] = tg_b(op == OP_MULT); r[40] = tg_b(op == OP_MULTU); r[41] = KB_ONE;
  if (hi_real) {
    const u32 value = e[9], rshard = e[10], ts = e[11], prev_value = e[12], prev_shard = e[13], prev_ts = e[14];
    tg_word(r + 42, prev_value);
    tg_word(r + 46, value);
    r[50] = tg_f(prev_shard); r[51] = tg_f(prev_ts);
    const bool same = prev_shard == rshard;
    r[52] = tg_b(same);
    const u32 d = (same ? ts - prev_ts : rshard - prev_shard) - 1u;
    r[53] = tg_f(d & 0xffffu); r[54] = tg_f((d >> 16) & 0xffu);
  } else {
    for (int i = 42; i < 55; i++) r[i] = 0;
  }
  r[55] = tg_b(hi_real);
  r[56] = hi_real ? tg_f(shard) : 0u; r[57] = hi_real ? tg_f(clk) : 0u;
}

// MemoryInstructionsChip::event_to_row, crates/core/machine/src/memory/instructions/trace.rs:103-263 (C++ twin
// include/memory_instrs.hpp).  Event: MemInstrEvent (crates/core/executor/src/events/instr.rs:114-136) as its 16 #[repr(C)]
// words {shard, clk, pc, next_pc, opcode, a, b, c, mem_access{tag, record}, prev_a_val}; the record is a MemoryReadRecord
// {value, shard, timestamp, prev_shard, prev_timestamp, -} for tag 0 and a MemoryWriteRecord {value, shard, timestamp,
// prev_value, prev_shard, prev_timestamp} for tag 1 (events/memory.rs:46-97).  Columns (79, columns.rs:15-117): pc, next_pc,
// shard, clk, op_a[4], op_b[4], op_c[4], is_lb .. is_sc (14), addr_word[4], addr_aligned, addr_ls_two_bits, ls_bits_is_one /
// two / three, addr_word_range_checker[14], memory_access {prev_value[4], value[4], prev_shard, prev_clk, compare_clk,
// diff_16bit_limb, diff_8bit_limb}, prev_a_val[4], unsigned_mem_val[4], most_sig_bit, most_sig_byte, mem_value_is_neg,
// most_sig_bytes_zero {inverse, result}.  The loaded value is one shift-and-mask per opcode family instead of byte arrays.
constexpr int MEMINSTR_WIDTH = 79;
enum : u32 { OP_LB = 31, OP_LBU = 32, OP_LH = 33, OP_LHU = 34, OP_LW = 35, OP_LWL = 36, OP_LWR = 37, OP_LL = 38, OP_SB = 39 };
KB_HD void fill_mem_instr(const u32* e, u32* r) {
  const u32 shard = e[0], clk = e[1], pc = e[2], next_pc = e[3], op = e[4] & 0xffu, a = e[5], b = e[6], c = e[7];
  const bool is_write = e[8] != 0;
  const u32 value = e[9], rshard = e[10], ts = e[11];
  const u32 prev_value = is_write ? e[12] : value, prev_shard = is_write ? e[13] : e[12], prev_ts = is_write ? e[14] : e[13];
  const u32 prev_a = e[15];
  r[0] = tg_f(pc); r[1] = tg_f(next_pc); r[2] = tg_f(shard); r[3] = tg_f(clk);
  tg_word(r + 4, a); tg_word(r + 8, b); tg_word(r + 12, c);
  tg_onehot(r + 16, 14, op - OP_LB);
  const u32 addr = b + c, ls = addr & 3u, sh = 8u * ls;
  tg_word(r + 30, addr);
  r[34] = tg_f(addr - ls); r[35] = tg_f(ls);
  r[36] = tg_b(ls == 1); r[37] = tg_b(ls == 2); r[38] = tg_b(ls == 3);
  tg_range_checker(r + 39, addr);
  tg_word(r + 53, prev_value); tg_word(r + 57, value);
  r[61] = tg_f(prev_shard); r[62] = tg_f(prev_ts);
  const bool same = prev_shard == rshard;
  r[63] = tg_b(same);
  const u32 d = (same ? ts - prev_ts : rshard - prev_shard) - 1u;
  r[64] = tg_f(d & 0xffffu); r[65] = tg_f((d >> 16) & 0xffu);
  tg_word(r + 66, prev_a);
  u32 um = 0;
  if (op == OP_LB || op == OP_LBU) um = (value >> sh) & 0xffu;
  else if (op == OP_LH || op == OP_LHU) um = (value >> (8u * (ls & 2u))) & 0xffffu;
  else if (op == OP_LW || op == OP_LL) um = value;
  else if (op == OP_LWL) um = (prev_a & ~(0xffffffffu << (24u - sh))) | (value << (24u - sh));
  else if (op == OP_LWR) um = (prev_a & ~(0xffffffffu >> sh)) | (value >> sh);
  tg_word(r + 70, um);
  const u32 msbyte = op == OP_LB ? um & 0xffu : op == OP_LH ? (um >> 8) & 0xffu : 0u;
  r[74] = tg_b(msbyte >> 7); r[75] = tg_f(msbyte); r[76] = tg_b(msbyte >> 7);
  const u32 upper = ((addr >> 8) & 0xffu) + ((addr >> 16) & 0xffu) + (addr >> 24);
  r[77] = upper ? fp_inv(fp_from_canonical(upper)).v : 0u;
  r[78] = tg_b(upper == 0);
}

// MemoryLocalChip::generate_trace, crates/core/machine/src/memory/local.rs:146-190 (C++ twin include/memory_local.hpp, one
// entry).  Event: MemoryLocalEvent (crates/core/executor/src/events/memory.rs:228-237) as its seven #[repr(C)] words {addr,
// initial_mem_access{shard, timestamp, value}, final_mem_access{shard, timestamp, value}}; FOUR events per row
// (NUM_LOCAL_MEMORY_ENTRIES_PER_ROW), event 4 i + k in entry k of row i.  Entry columns (14, local.rs:29-55): addr,
// initial_shard, final_shard, initial_clk, final_clk, initial_value[4], final_value[4], is_real; entries past the last event
// are zero.
constexpr int MEMLOCAL_ENTRIES = 4, MEMLOCAL_ENTRY_WIDTH = 14;
KB_HD void fill_memory_local(const u32* e, int n_valid, u32* r) {
#pragma unroll
  for (int k = 0; k < MEMLOCAL_ENTRIES; k++, e += 7, r += MEMLOCAL_ENTRY_WIDTH) {
    if (k < n_valid) {
      r[0] = tg_f(e[0]);
      r[1] = tg_f(e[1]); r[2] = tg_f(e[4]);
      r[3] = tg_f(e[2]); r[4] = tg_f(e[5]);
      tg_word(r + 5, e[3]); tg_word(r + 9, e[6]);
      r[13] = KB_ONE;
    } else {
      for (int i = 0; i < MEMLOCAL_ENTRY_WIDTH; i++) r[i] = 0;
    }
  }
}

// CpuChip::event_to_row, crates/core/machine/src/cpu/trace.rs:119-246 (C++ twin include/cpu.hpp).  CpuEvent holds Options
// and the instruction is fetched from the program, so the host flattens both into a 28-word `zkb200_cpu_event`
// (include/zkb200.h): clk, pc, next_pc, next_next_pc, a, b, c, hi, flags (bit 0 hi is Some; bits 1-2 a_record: 0 None, 1 Read,
// 2 Write; bit 3 b_record is Read; bit 4 c_record is Read; bit 5 imm_b; bit 6 imm_c), opcode | op_a << 8 | shard << 16, op_b,
// op_c, a_record (six words laid out as the MemoryRecordEnum payload: read {value, shard, timestamp, prev_shard,
// prev_timestamp, -}, write {value, shard, timestamp, prev_value, prev_shard, prev_timestamp}), b_record and c_record
// (MemoryReadRecord, five words each).  Columns (67, cpu/columns/mod.rs:18-84): shard, clk_16bit_limb, clk_8bit_limb,
// shard_to_send, clk_to_send, pc, next_pc, next_next_pc, instruction {opcode, op_a, op_b[4], op_c[4], op_a_0, imm_b, imm_c},
// num_extra_cycles, is_rw_a, is_check_memory, is_halt, is_sequential, op_a_value[4], hi_or_prev_a[4], op_a_access
// {prev_value[4], value[4], prev_shard, prev_clk, compare_clk, diff_16bit_limb, diff_8bit_limb}, op_b_access[9],
// op_c_access[9], is_real, op_a_immutable.  The instruction classes (crates/core/executor/src/instruction.rs:147-310) are bit
// masks over the opcode number.
constexpr int CPU_WIDTH = 67, CPU_EVENT_WORDS = 28;
KB_HD constexpr u64 tg_opmask(int lo, int hi) { return ((~0ull) >> (63 - (hi - lo))) << lo; }     // opcodes lo..hi
constexpr u64 CPU_M_LOADS = tg_opmask(31, 38), CPU_M_STORES_NO_SC = tg_opmask(39, 43), CPU_M_SC = 1ull << 44,
              CPU_M_BRANCH = tg_opmask(21, 26), CPU_M_JUMP = tg_opmask(27, 29), CPU_M_SYSCALL = 1ull << 30,
              CPU_M_MADDSUB = tg_opmask(46, 49), CPU_M_INS = 1ull << 45, CPU_M_MOVCOND = tg_opmask(50, 51), CPU_M_TEQ = 1ull << 54,
              CPU_M_MULT_DIV = tg_opmask(3, 6);
constexpr u64 CPU_M_CHECK_MEMORY = CPU_M_SYSCALL | CPU_M_MADDSUB | CPU_M_LOADS | CPU_M_STORES_NO_SC | CPU_M_SC;
constexpr u64 CPU_M_RW_A = CPU_M_CHECK_MEMORY | CPU_M_INS | CPU_M_MOVCOND;
// MemoryAccessCols::populate_access (memory/consistency/trace.rs:56-98): prev_shard, prev_clk, compare_clk and the two limbs
// of (current - previous - 1) in the compared time
KB_HD void tg_mem_access_tail(u32* r, u32 shard, u32 ts, u32 prev_shard, u32 prev_ts) {
  r[0] = tg_f(prev_shard); r[1] = tg_f(prev_ts);
  const bool same = prev_shard == shard;
  r[2] = tg_b(same);
  const u32 d = (same ? ts - prev_ts : shard - prev_shard) - 1u;
  r[3] = tg_f(d & 0xffffu); r[4] = tg_f((d >> 16) & 0xffu);
}
KB_HD void fill_cpu(const u32* e, u32* r) {
  const u32 clk = e[0], pc = e[1], next_pc = e[2], next_next_pc = e[3], a = e[4], b = e[5], c = e[6], hi = e[7], fl = e[8];
  const u32 op = e[9] & 0xffu, op_a = (e[9] >> 8) & 0xffu, shard = e[9] >> 16, op_b = e[10], op_c = e[11];
  const u64 bit = op < 64 ? 1ull << op : 0ull;
  const bool check_or_multdiv = (bit & (CPU_M_CHECK_MEMORY | CPU_M_MULT_DIV)) != 0;
  r[0] = tg_f(shard); r[1] = tg_f(clk & 0xffffu); r[2] = tg_f((clk >> 16) & 0xffu);
  r[3] = check_or_multdiv ? tg_f(shard) : 0u; r[4] = check_or_multdiv ? tg_f(clk) : 0u;
  r[5] = tg_f(pc); r[6] = tg_f(next_pc); r[7] = tg_f(next_next_pc);
  r[8] = tg_f(op); r[9] = tg_f(op_a);
  tg_word(r + 10, op_b); tg_word(r + 14, op_c);
  r[18] = tg_b(op_a == 0); r[19] = tg_b((fl >> 5) & 1u); r[20] = tg_b((fl >> 6) & 1u);
  r[22] = tg_b((bit & CPU_M_RW_A) != 0);
  r[23] = tg_b(check_or_multdiv);
  tg_word(r + 26, a);
  tg_word(r + 30, (fl & 1u) ? hi : 0u);
  // op_a_access: the value word is `a` unless a record overwrites it
  const u32 a_kind = (fl >> 1) & 3u;
  const u32* ra = e + 12;
  u32 a_prev_value = 0;
  if (a_kind == 0) {
    for (int i = 34; i < 38; i++) r[i] = 0;
    tg_word(r + 38, a);
    for (int i = 42; i < 47; i++) r[i] = 0;
  } else {
    const bool wr = a_kind == 2;
    a_prev_value = wr ? ra[3] : ra[0];
    tg_word(r + 34, a_prev_value);
    tg_word(r + 38, ra[0]);
    tg_mem_access_tail(r + 42, ra[1], ra[2], wr ? ra[4] : ra[3], wr ? ra[5] : ra[4]);
  }
#pragma unroll
  for (int k = 0; k < 2; k++) {                       // op_b_access, op_c_access
    u32* rc = r + 47 + 9 * k;
    const u32* rec = e + 18 + 5 * k;
    if ((fl >> (3 + k)) & 1u) {
      tg_word(rc, rec[0]);
      tg_mem_access_tail(rc + 4, rec[1], rec[2], rec[3], rec[4]);
    } else {
      tg_word(rc, k ? c : b);
      for (int i = 4; i < 9; i++) rc[i] = 0;
    }
  }
  // SYSCALL: HALT (id 0) and SYS_EXT_GROUP (4246) end the program; the syscall code is register a's previous value
  bool is_halt = false;
  u32 extra = 0;
  if (bit & CPU_M_SYSCALL) {
    const u32 id = a_prev_value & 0xffffu;
    is_halt = id == 0u || id == (4246u & 0xffffu);
    extra = tg_f(a_prev_value >> 24);
  }
  r[21] = extra; r[24] = tg_b(is_halt);
  r[25] = tg_b(!is_halt && !(bit & (CPU_M_BRANCH | CPU_M_JUMP)));
  r[65] = KB_ONE;
  r[66] = tg_b((bit & (CPU_M_STORES_NO_SC | CPU_M_BRANCH | CPU_M_TEQ)) != 0);
}

// MiscInstrsChip::event_to_row, crates/core/machine/src/misc/others/trace.rs:91-275 (C++ twin include/misc_instrs.hpp).  Event:
// MiscEvent (crates/core/executor/src/events/instr.rs:241-261) as its 15 #[repr(C)] words {shard, clk, pc, next_pc, opcode, a,
// b, c, prev_a, hi_record{value, shard, timestamp, prev_value, prev_shard, prev_timestamp}}.  Columns (72,
// misc/others/columns/mod.rs): shard, clk, pc, next_pc, op_a[4], prev_a[4], op_b[4], op_c[4], a 44-column UNION viewed per
// opcode family, is_sext, is_ins, is_ext, is_maddu, is_msubu, is_madd, is_msub, is_teq.  Union views (columns/*.rs):
//   maddsub  mul_lo[4], mul_hi[4], add_operation{value[4], value_hi[4], carry[7]}, src2_hi[4], src2_lo[4], op_hi_access[13]
//   sext     most_sig_bit, sig_byte, a_eq_b{is_zero_byte[4]{inverse, result}, lower_half_zero, upper_half_zero, result}, is_seb, is_seh
//   ext      lsb, msbd, sll_val[4]            ins   lsb, msb, ror_val[4], srl1_val[4], srl_val[4], sll_val[4], add_val[4]
constexpr int MISC_WIDTH = 72, MISC_EVENT_WORDS = 15;
enum : u32 { OP_INS = 45, OP_MADDU = 46, OP_MSUBU = 47, OP_MADD = 48, OP_MSUB = 49, OP_EXT = 53, OP_TEQ = 54, OP_SEXT = 55 };
KB_HD void fill_misc(const u32* e, u32* r, const u32* inv255) {
  const u32 shard = e[0], clk = e[1], pc = e[2], next_pc = e[3], op = e[4] & 0xffu, a = e[5], b = e[6], c = e[7], prev_a = e[8];
  r[0] = tg_f(shard); r[1] = tg_f(clk); r[2] = tg_f(pc); r[3] = tg_f(next_pc);
  tg_word(r + 4, a); tg_word(r + 8, prev_a); tg_word(r + 12, b); tg_word(r + 16, c);
  u32* u = r + 20;
  for (int i = 0; i < 44; i++) u[i] = 0;
  r[64] = tg_b(op == OP_SEXT); r[65] = tg_b(op == OP_INS); r[66] = tg_b(op == OP_EXT); r[67] = tg_b(op == OP_MADDU);
  r[68] = tg_b(op == OP_MSUBU); r[69] = tg_b(op == OP_MADD); r[70] = tg_b(op == OP_MSUB); r[71] = tg_b(op == OP_TEQ);
  if (op == OP_SEXT || op == OP_TEQ) {
    const bool seh = c > 0;
    const u32 sig_byte = seh ? (b >> 8) & 0xffu : b & 0xffu;
    u[0] = tg_b(sig_byte >> 7); u[1] = tg_f(sig_byte);
    u32 zero_mask = 0;
    for (int i = 0; i < 4; i++) {
      const u32 x = (a >> (8 * i)) & 0xffu, y = (b >> (8 * i)) & 0xffu;
      // 1 / (x - y): the table holds 1/d for d = 1..255, a negative difference negates it
      u[2 + 2 * i] = x > y ? inv255[x - y] : x < y ? KB_P - inv255[y - x] : 0u;
      u[3 + 2 * i] = tg_b(x == y);
      zero_mask |= (u32)(x == y) << i;
    }
    u[10] = tg_b((zero_mask & 3u) == 3u); u[11] = tg_b((zero_mask & 12u) == 12u); u[12] = tg_b(zero_mask == 15u);
    u[13] = tg_b(!seh); u[14] = tg_b(seh);
  } else if (op >= OP_MADDU && op <= OP_MSUB) {
    const bool is_sign = op == OP_MADD || op == OP_MSUB, is_add = op == OP_MADDU || op == OP_MADD;
    const u64 multiply = is_sign ? (u64)((int64_t)(int32_t)b * (int64_t)(int32_t)c) : (u64)b * (u64)c;
    const u32 hv = e[9], hshard = e[10], hts = e[11], hprev_value = e[12], hprev_shard = e[13], hprev_ts = e[14];
    tg_word(u + 0, (u32)multiply); tg_word(u + 4, (u32)(multiply >> 32));
    const u32 src2_lo = is_add ? prev_a : a, src2_hi = is_add ? hprev_value : hv;
    const u64 src2 = ((u64)src2_hi << 32) + src2_lo, sum = multiply + src2;
    tg_long(u + 8, sum);
#pragma unroll
    for (int k = 0; k < 7; k++) {                   // carry out of the low k + 1 bytes
      const u64 mask = (1ull << (8 * (k + 1))) - 1ull;
      u[16 + k] = tg_b((((multiply & mask) + (src2 & mask)) >> (8 * (k + 1))) != 0);
    }
    tg_word(u + 23, src2_hi); tg_word(u + 27, src2_lo);
    tg_word(u + 31, hprev_value); tg_word(u + 35, hv);
    u[39] = tg_f(hprev_shard); u[40] = tg_f(hprev_ts);
    const bool same = hprev_shard == hshard;
    u[41] = tg_b(same);
    const u32 d = (same ? hts - hprev_ts : hshard - hprev_shard) - 1u;
    u[42] = tg_f(d & 0xffffu); u[43] = tg_f((d >> 16) & 0xffu);
  } else if (op == OP_EXT) {
    const u32 lsb = c & 0x1fu, msbd = c >> 5;
    u[0] = tg_f(lsb); u[1] = tg_f(msbd);
    tg_word(u + 2, b << ((31u - lsb - msbd) & 31u));
  } else if (op == OP_INS) {
    const u32 lsb = c & 0x1fu, msb = c >> 5;
    const u32 ror_val = lsb ? (prev_a >> lsb) | (prev_a << (32u - lsb)) : prev_a;
    const u32 srl1_val = ror_val >> 1, srl_val = srl1_val >> ((msb - lsb) & 31u), sll_val = b << ((31u - msb + lsb) & 31u);
    u[0] = tg_f(lsb); u[1] = tg_f(msb);
    tg_word(u + 2, ror_val); tg_word(u + 6, srl1_val); tg_word(u + 10, srl_val); tg_word(u + 14, sll_val);
    tg_word(u + 18, srl_val + sll_val);
  }
}

// IsEqualWordOperation::populate (operations/is_equal_word.rs) = IsZeroWordOperation over the byte differences x[i] - y[i] as
// field elements (is_zero_word.rs: is_zero_byte[4]{inverse, result}, is_lower_half_zero, is_upper_half_zero, result); y = 0
// gives IsZeroWordOperation::populate(x).  Eleven columns.
KB_HD void tg_is_equal_word(u32* r, u32 x, u32 y, const u32* inv255) {
  u32 zero_mask = 0;
  for (int i = 0; i < 4; i++) {
    const u32 p = (x >> (8 * i)) & 0xffu, q = (y >> (8 * i)) & 0xffu;
    r[2 * i] = p > q ? inv255[p - q] : p < q ? KB_P - inv255[q - p] : 0u;
    r[2 * i + 1] = tg_b(p == q);
    zero_mask |= (u32)(p == q) << i;
  }
  r[8] = tg_b((zero_mask & 3u) == 3u); r[9] = tg_b((zero_mask & 12u) == 12u); r[10] = tg_b(zero_mask == 15u);
}
// MemoryReadWriteCols::populate of a MemoryWriteRecord {value, shard, timestamp, prev_value, prev_shard, prev_timestamp}
// (crates/core/machine/src/memory/trace.rs): prev_value[4], value[4], prev_shard, prev_clk, compare_clk, diff_16bit_limb,
// diff_8bit_limb.  Thirteen columns.
KB_HD void tg_write_access(u32* r, const u32* rec) {
  tg_word(r, rec[3]); tg_word(r + 4, rec[0]);
  r[8] = tg_f(rec[4]); r[9] = tg_f(rec[5]);
  const bool same = rec[4] == rec[1];
  r[10] = tg_b(same);
  const u32 d = (same ? rec[2] - rec[5] : rec[1] - rec[4]) - 1u;
  r[11] = tg_f(d & 0xffffu); r[12] = tg_f((d >> 16) & 0xffu);
}
// IsZeroOperation::populate_from_field_element (operations/is_zero.rs:29-40) of the difference x - k of two small integers
KB_HD void tg_is_zero_diff(u32* r, u32 x, u32 k) {
  r[0] = x == k ? 0u : fp_inv(fp_from_canonical(x) - fp_from_canonical(k)).v;
  r[1] = tg_b(x == k);
}

// DivRemChip::generate_trace, crates/core/machine/src/alu/divrem/mod.rs:229-364 (C++ twin include/div_rem.hpp, which differs
// from it for c = 0 and for INT_MIN / -1; the Rust is followed).  Event: CompAluEvent as for Mul.  Columns (106,
// divrem/mod.rs:109-204): pc, next_pc, b[4], c[4], quotient[4], remainder[4], abs_remainder[4], abs_c[4], max_abs_c_or_1[4],
// c_times_quotient[8], carry[8], is_c_0[11], is_div, is_divu, is_mod, is_modu, is_overflow, is_overflow_b[11],
// is_overflow_c[11], b_msb, rem_msb, c_msb, b_neg, rem_neg, c_neg, remainder_check_multiplicity, op_hi_access[13], shard, clk.
constexpr int DIVREM_WIDTH = 106;
enum : u32 { OP_DIV = 5, OP_DIVU = 6, OP_MOD = 7, OP_MODU = 8 };
KB_HD void fill_div_rem(const u32* e, u32* r, const u32* inv255) {
  const u32 shard = e[0], clk = e[1], pc = e[2], next_pc = e[3], op = e[4] & 0xffu, b = e[7], c = e[8];
  const bool sgn = op == OP_DIV || op == OP_MOD, overflow = b == 0x80000000u && c == 0xffffffffu;
  // get_quotient_and_remainder, crates/core/executor/src/utils.rs:33-43 (wrapping division)
  u32 quot, rem;
  if (c == 0) { quot = 0xffffffffu; rem = b; }
  else if (!sgn) { quot = b / c; rem = b % c; }
  else if (overflow) { quot = 0x80000000u; rem = 0; }
  else { quot = (u32)((int32_t)b / (int32_t)c); rem = (u32)((int32_t)b % (int32_t)c); }
  r[0] = tg_f(pc); r[1] = tg_f(next_pc);
  tg_word(r + 2, b); tg_word(r + 6, c); tg_word(r + 10, quot); tg_word(r + 14, rem);
  const u32 abs_rem = sgn && (rem >> 31) ? 0u - rem : rem, abs_c = sgn && (c >> 31) ? 0u - c : c;
  tg_word(r + 18, abs_rem); tg_word(r + 22, abs_c); tg_word(r + 26, abs_c ? abs_c : 1u);
  const u64 ctq = sgn ? (u64)((int64_t)(int32_t)quot * (int64_t)(int32_t)c) : (u64)quot * (u64)c;
  const u64 remw = sgn ? (u64)(int64_t)(int32_t)rem : (u64)rem;
  u32 carry = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r[30 + i] = tg_f((u32)(ctq >> (8 * i)) & 0xffu);
    carry = ((u32)((ctq >> (8 * i)) & 0xffu) + (u32)((remw >> (8 * i)) & 0xffu) + carry) >> 8;
    r[38 + i] = tg_b(carry);
  }
  tg_is_equal_word(r + 46, c, 0u, inv255);
  r[57] = tg_b(op == OP_DIV); r[58] = tg_b(op == OP_DIVU); r[59] = tg_b(op == OP_MOD); r[60] = tg_b(op == OP_MODU);
  r[61] = tg_b(sgn && overflow);
  tg_is_equal_word(r + 62, b, 0x80000000u, inv255);
  tg_is_equal_word(r + 73, c, 0xffffffffu, inv255);
  r[84] = tg_b(b >> 31); r[85] = tg_b(rem >> 31); r[86] = tg_b(c >> 31);
  r[87] = tg_b(sgn && (b >> 31)); r[88] = tg_b(sgn && (rem >> 31)); r[89] = tg_b(sgn && (c >> 31));
  r[90] = tg_b(c != 0);
  if (op == OP_DIV || op == OP_DIVU) {          // the HI write of DIV / DIVU (mod.rs:259-267)
    tg_write_access(r + 91, e + 9);
    r[104] = tg_f(shard); r[105] = tg_f(clk);
  } else {
    for (int i = 91; i < 106; i++) r[i] = 0;
  }
}

// SyscallChip::generate_trace row_fn, crates/core/machine/src/syscall/chip.rs:189-230 (C++ twin include/syscall.hpp).  Event:
// SyscallEvent (crates/core/executor/src/events/syscall.rs:8-29) as its 14 #[repr(C)] words {pc, next_pc, shard, clk,
// a_record{value, shard, timestamp, prev_value, prev_shard, prev_timestamp}, a_record_is_real, syscall_id, arg1, arg2}.
// SyscallCore takes the events that are sent to the table (prev_value byte 2 = 1 or byte 1 != 0, chip.rs:233-240: the caller
// filters); SyscallPrecompile takes one event per precompile event with a_record.prev_value = 1 / a_record.value = v0 for
// PrecompileEvent::Linux and prev_value = 0 otherwise (the convention of syscall.hpp precompile_event_to_row).
// Columns (11): shard, clk, syscall_id, arg1_lo, arg1_hi, arg2_lo, arg2_hi, result_lo, result_hi, is_linux, is_real.
constexpr int SYSCALL_WIDTH = 11, SYSCALL_EVENT_WORDS = 14;
KB_HD void fill_syscall(const u32* e, u32* r, bool precompile) {
  const u32 value = e[4], prev_value = e[7], arg1 = e[12], arg2 = e[13];
  const bool is_linux = precompile ? prev_value == 1u : ((prev_value >> 8) & 0xffu) != 0;
  r[0] = tg_f(e[2]); r[1] = tg_f(e[3]); r[2] = tg_f(e[11]);
  r[3] = tg_f(arg1 & 0xffffu); r[4] = tg_f(arg1 >> 16); r[5] = tg_f(arg2 & 0xffffu); r[6] = tg_f(arg2 >> 16);
  r[7] = is_linux ? tg_f(value & 0xffffu) : 0u; r[8] = is_linux ? tg_f(value >> 16) : 0u;
  r[9] = tg_b(is_linux); r[10] = KB_ONE;
}

// SyscallInstrsChip::event_to_row, crates/core/machine/src/syscall/instructions/trace.rs:89-177 (C++ twin
// include/syscall_instrs.hpp).  Event: SyscallEvent.  Columns (77, instructions/columns.rs:11-59): pc, next_pc, shard, clk,
// num_extra_cycles, is_halt, is_sys_linux, is_prev_a1_zero{inverse, result}, syscall_id, op_a[4], op_b[4], op_c[4], prev_a[4],
// is_enter_unconstrained, is_hint_len, is_halt_check, is_exit_group_check, is_commit, is_commit_deferred_proofs (each
// {inverse, result}), index_bitmap[8], op_b_range_check[14], op_c_range_check[14], op_b_check, op_c_check, is_real.
constexpr int SYSINSTR_WIDTH = 77;
enum : u32 { SYS_HALT = 0x00, SYS_ENTER_UNCONSTRAINED = 0x03, SYS_COMMIT = 0x10, SYS_COMMIT_DEFERRED_PROOFS = 0x1a, SYS_HINT_LEN = 0xf0,
             SYS_EXT_GROUP = 4246 };           // SyscallCode::syscall_id(), crates/core/executor/src/syscalls/code.rs
KB_HD void fill_syscall_instr(const u32* e, u32* r, const u32* inv255) {
  const u32 value = e[4], prev_value = e[7], arg1 = e[12], arg2 = e[13];
  const u32 sid = prev_value & 0xffffu, b1 = (prev_value >> 8) & 0xffu, b2 = (prev_value >> 16) & 0xffu;
  const bool is_halt = sid == SYS_HALT || sid == SYS_EXT_GROUP, send_to_table = b1 != 0 || b2 == 1;
  r[0] = tg_f(e[0]); r[1] = tg_f(e[1]); r[2] = tg_f(e[2]); r[3] = tg_f(e[3]);
  r[4] = tg_f(prev_value >> 24);
  r[5] = tg_b(is_halt); r[6] = tg_b(b1 != 0);
  r[7] = b1 ? inv255[b1] : 0u; r[8] = tg_b(b1 == 0);
  r[9] = tg_f(e[11]);
  tg_word(r + 10, value); tg_word(r + 14, arg1); tg_word(r + 18, arg2); tg_word(r + 22, prev_value);
  tg_is_zero_diff(r + 26, sid, SYS_ENTER_UNCONSTRAINED); tg_is_zero_diff(r + 28, sid, SYS_HINT_LEN);
  tg_is_zero_diff(r + 30, sid, SYS_HALT); tg_is_zero_diff(r + 32, sid, SYS_EXT_GROUP);
  tg_is_zero_diff(r + 34, sid, SYS_COMMIT); tg_is_zero_diff(r + 36, sid, SYS_COMMIT_DEFERRED_PROOFS);
  const bool commits = sid == SYS_COMMIT || sid == SYS_COMMIT_DEFERRED_PROOFS;
  for (u32 i = 0; i < 8; i++) r[38 + i] = tg_b(commits && arg1 == i);
  const bool b_check = send_to_table || is_halt, c_check = send_to_table || sid == SYS_COMMIT_DEFERRED_PROOFS;
  if (b_check) tg_range_checker(r + 46, arg1); else for (int i = 46; i < 60; i++) r[i] = 0;
  if (c_check) tg_range_checker(r + 60, arg2); else for (int i = 60; i < 74; i++) r[i] = 0;
  r[74] = tg_b(b_check); r[75] = tg_b(c_check); r[76] = KB_ONE;
}

// MemoryGlobalChip::generate_trace, crates/core/machine/src/memory/global.rs:115-192 (C++ twin include/memory_global.hpp for
// the columns that depend on the event alone).  Event: zkb200_memory_global_event, the MemoryInitializeFinalizeEvent {addr,
// value, shard, timestamp} (events/memory.rs:138-149) of the address-sorted vector followed by prev_addr - the previous event's
// address, for the first event the public values' previous_init / previous_finalize address - and position (bit 0: first
// event, bit 1: last event), the two things the reference's sequential second loop reads from the neighbouring row.
// Columns (111, global.rs:210-245): shard, timestamp, addr, lt_cols.bit_flags[32], addr_bits{bits[32],
// and_most_sig_byte_decomp_0_to_2 .. 0_to_7}, value[32], is_real, is_next_comp, is_prev_addr_zero{inverse, result},
// is_first_comp, is_last_addr.
constexpr int MEMGLOBAL_WIDTH = 111, MEMGLOBAL_EVENT_WORDS = 6;
KB_HD void fill_memory_global(const u32* e, u32* r) {
  const u32 addr = e[0], value = e[1], prev_addr = e[4];
  const bool first = e[5] & 1u, last = (e[5] >> 1) & 1u;
  r[0] = tg_f(e[2]); r[1] = tg_f(e[3]); r[2] = tg_f(addr);
  // AssertLtColsBits::populate (operations/cmp.rs:300-319): the most significant bit where prev_addr < addr differ
  const bool compare = (!first || prev_addr != 0) && prev_addr != addr;
  tg_onehot(r + 3, 32, compare ? tg_top_bit(prev_addr ^ addr) : 32u);
  tg_bits(r + 35, 32, addr);
  u32 acc = (addr >> 24) & (addr >> 25) & 1u;
  r[67] = tg_b(acc);
  for (int i = 2; i <= 6; i++) { acc &= addr >> (24 + i); r[66 + i] = tg_b(acc & 1u); }
  tg_bits(r + 73, 32, value);
  r[105] = KB_ONE; r[106] = tg_b(!first);
  r[107] = first && prev_addr ? fp_inv(fp_from_canonical(prev_addr)).v : 0u;
  r[108] = tg_b(first && prev_addr == 0);
  r[109] = tg_b(first && prev_addr != 0);
  r[110] = tg_b(last);
}

// Rows past the last event (generate_trace of each chip): zeros, except the dummy rows that keep the
// shift and count-leading chips' constraints satisfied (sll/mod.rs:160-173, sr/mod.rs:184-187,
// clo_clz/mod.rs:150-163).
KB_HD void fill_alu_padding(int chip, u32* r) {
  const int w = alu_width(chip);
  for (int i = 0; i < w; i++) r[i] = 0;
  if (chip == ALU_SLL) { r[22] = KB_ONE; r[30] = KB_ONE; r[39] = KB_ONE; }
  if (chip == ALU_SR) { r[10] = KB_ONE; r[18] = KB_ONE; }
  if (chip == ALU_CLOCLZ) { tg_word(r + 2, 32); r[14] = KB_ONE; }
  if (chip == ALU_CPU) { r[19] = KB_ONE; r[20] = KB_ONE; r[22] = KB_ONE; }      // imm_b, imm_c, is_rw_a (cpu/trace.rs:60-63)
}

// w: the row's event words as they lie in the record's event vector (alu_events_per_row x alu_event_words of them, the first
// n_valid events present)
KB_HD void fill_alu_row(int chip, const u32* w, u32* r, const u32* inv255, int n_valid = 1) {
  switch (chip) {
    case ALU_ADDSUB: fill_add_sub(alu_event_from_words(w), r); break;
    case ALU_BITWISE: fill_bitwise(alu_event_from_words(w), r); break;
    case ALU_LT: fill_lt(alu_event_from_words(w), r, inv255); break;
    case ALU_SLL: fill_shift_left(alu_event_from_words(w), r); break;
    case ALU_SR: fill_shift_right(alu_event_from_words(w), r); break;
    case ALU_CLOCLZ: fill_clo_clz(alu_event_from_words(w), r); break;
    case ALU_BRANCH: fill_branch(flow_event_from_words(w), r); break;
    case ALU_JUMP: fill_jump(flow_event_from_words(w), r); break;
    case ALU_MUL: fill_mul(w, r); break;
    case ALU_MEMINSTR: fill_mem_instr(w, r); break;
    case ALU_MEMLOCAL: fill_memory_local(w, n_valid, r); break;
    case ALU_CPU: fill_cpu(w, r); break;
    case ALU_MISC: fill_misc(w, r, inv255); break;
    case ALU_DIVREM: fill_div_rem(w, r, inv255); break;
    case ALU_SYSCALL_CORE: fill_syscall(w, r, false); break;
    case ALU_SYSCALL_PRECOMPILE: fill_syscall(w, r, true); break;
    case ALU_SYSCALL_INSTRS: fill_syscall_instr(w, r, inv255); break;
    case ALU_MEMGLOBAL_INIT: case ALU_MEMGLOBAL_FINALIZE: fill_memory_global(w, r); break;
    default: fill_mov_cond(w, r, inv255); break;
  }
}

// One CTA of the row kernel (csrc/tracegen.cu alu_rows_kernel), as three phases separated by CTA barriers: one thread per row,
// R rows per CTA.  Written host/device so that tests/hostcheck can walk the same index arithmetic thread by thread.
//   load   the CTA's events are at most R * RW consecutive words: coalesced copy into shared memory
//   fill   one row per thread into a tile whose row stride WP is odd (conflict-free column-wise read-back)
//   store  the tile with coalesced writes: R rows are one contiguous block of the row-major matrix, and R consecutive
//          elements of every column of the column-major one
// R = 128, or 64 for the chips whose events and tile would not fit the 48 KB of static shared memory (DivRem, MemoryGlobal*).
KB_HD constexpr int alu_cta_rows(int chip) {
  return (size_t)128 * (size_t)(alu_event_words(chip) * alu_events_per_row(chip) + (alu_width(chip) | 1)) * sizeof(u32) <= 48 * 1024 ? 128 : 64;
}
template <int CHIP>
struct AluCta {
  static constexpr int W = alu_width(CHIP), WP = W | 1, EW = alu_event_words(CHIP), EPR = alu_events_per_row(CHIP);
  static constexpr int RW = EW * EPR;                  // event words per row
  static constexpr int R = alu_cta_rows(CHIP);         // rows per CTA = threads per CTA
  static_assert((size_t)R * (RW + WP) * sizeof(u32) <= 48 * 1024, "events and row tile must fit the static shared-memory limit");
  static KB_HD void load(u32 tid, size_t cta, const u32* events, size_t n, u32* ev_s) {
    const size_t e0 = cta * R * EPR;
    const size_t ev_words = e0 < n ? (n - e0 < (size_t)R * EPR ? (n - e0) * EW : (size_t)R * RW) : 0;
    for (u32 i = tid; i < ev_words; i += R) ev_s[i] = events[e0 * EW + i];
  }
  static KB_HD void fill(u32 tid, size_t cta, size_t n, const u32* ev_s, u32* tile, const u32* inv255) {
    u32* r = tile + tid * WP;
    const size_t first = (cta * R + tid) * EPR;        // the row's first event
    if (first < n) fill_alu_row(CHIP, ev_s + RW * tid, r, inv255, n - first < (size_t)EPR ? (int)(n - first) : EPR);
    else fill_alu_padding(CHIP, r);
  }
  static KB_HD void store(u32 tid, size_t cta, size_t height, const u32* tile, u32* out, int col_major) {
    const size_t row0 = cta * R;
    const size_t rows = height - row0 < (size_t)R ? height - row0 : (size_t)R;
    if (col_major) {
      if (tid < rows) {
#pragma unroll 4
        for (int c = 0; c < W; c++) out[(size_t)c * height + row0 + tid] = tile[tid * WP + c];
      }
    } else {
      u32* dst = out + row0 * W;
      for (u32 i = tid; i < rows * W; i += R) dst[i] = tile[(i / W) * WP + (i % W)];
    }
  }
};

// 1/d for d = 0..255 in Montgomery form (entry 0 unused)
inline void alu_build_inv255(u32* out) {
  out[0] = 0;
  for (u32 d = 1; d < 256; d++) out[d] = fp_inv(fp_from_canonical(d)).v;
}

}  // namespace zkb

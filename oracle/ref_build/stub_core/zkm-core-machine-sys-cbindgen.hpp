// Stand-in for the header cbindgen generates in the reference's Rust build
// (crates/core/machine/build.rs:94-180); that build cannot run here (no Rust toolchain).  It
// declares, with the field order of the Rust #[repr(C)] definitions cited below, the layout types
// the reference's ALU and control-flow row fillers (crates/core/machine/include/{add_sub,bitwise,lt,
// shift_left,shift_right,clo_clz,branch,jump,mov_cond}.hpp) name, plus the names utils.hpp mentions in signatures.  This file is
// ours; the reference's headers are compiled from where they lie under /root/reference.
#pragma once
#include <cstddef>
#include <cstdint>

namespace zkm_core_machine_sys {

constexpr size_t BYTE_SIZE = 8;        // crates/core/machine/src/alu/sll/mod.rs (BYTE_SIZE)
constexpr size_t WORD_SIZE = 4;        // crates/primitives/src/consts.rs:5
constexpr size_t LONG_WORD_SIZE = 8;   // crates/primitives/src/consts.rs:6
constexpr size_t PRODUCT_SIZE = 8;     // crates/core/machine/src/alu/mul/mod.rs:64

// crates/core/executor/src/opcode.rs:25-89 (#[repr(u8)])
enum class Opcode : uint8_t {
  ADD = 0, SUB = 1, MUL = 2, MULT = 3, MULTU = 4, DIV = 5, DIVU = 6, MOD = 7, MODU = 8, SLL = 9, SRL = 10, SRA = 11,
  ROR = 12, SLT = 13, SLTU = 14, AND = 15, OR = 16, XOR = 17, NOR = 18, CLZ = 19, CLO = 20,
  BEQ = 21, BGEZ = 22, BGTZ = 23, BLEZ = 24, BLTZ = 25, BNE = 26, Jump = 27, Jumpi = 28, JumpDirect = 29, SYSCALL = 30,
  LB = 31, LBU = 32, LH = 33, LHU = 34, LW = 35, LWL = 36, LWR = 37, LL = 38, SB = 39, SH = 40, SW = 41, SWL = 42,
  SWR = 43, SC = 44, INS = 45, MADDU = 46, MSUBU = 47, MADD = 48, MSUB = 49, MEQ = 50, MNE = 51, WSBH = 52, EXT = 53,
  TEQ = 54, SEXT = 55, UNIMPL = 0xff,
};
// crates/core/executor/src/syscalls/code.rs:28-154 (the codes cpu.hpp and syscall_instrs.hpp name; utils.hpp:to_syscall_id
// takes the type)
enum class SyscallCode : uint32_t {
  HALT = 0x00, ENTER_UNCONSTRAINED = 0x03, COMMIT = 0x10, COMMIT_DEFERRED_PROOFS = 0x1A, SYSHINTLEN = 0xF0, SYS_EXT_GROUP = 4246,
};
constexpr size_t PV_DIGEST_NUM_WORDS = 8;   // crates/stark/src/air/public_values.rs

// crates/core/executor/src/events/instr.rs:11-26 (#[repr(C)])
struct AluEvent {
  uint32_t pc;
  uint32_t next_pc;
  Opcode opcode;
  uint32_t hi;
  uint32_t a;
  uint32_t b;
  uint32_t c;
};

// crates/core/executor/src/events/instr.rs:160-217 (#[repr(C)])
struct BranchEvent { uint32_t pc, next_pc, next_next_pc; Opcode opcode; uint32_t a, b, c; };
struct JumpEvent { uint32_t pc, next_pc, next_next_pc; Opcode opcode; uint32_t a, b, c; };
// crates/core/executor/src/events/instr.rs:287-302 (#[repr(C)])
struct MovCondEvent { uint32_t pc, next_pc; Opcode opcode; uint32_t a, b, c, prev_a; };

template <class T> struct Word { T _0[WORD_SIZE]; };   // crates/stark/src/word.rs:21

// operation helpers utils.hpp takes by reference (crates/core/machine/src/operations/*.rs)
template <class T> struct KoalaBearWordRangeChecker {
  T most_sig_byte_decomp[8];
  T and_most_sig_byte_decomp_0_to_2, and_most_sig_byte_decomp_0_to_3, and_most_sig_byte_decomp_0_to_4,
      and_most_sig_byte_decomp_0_to_5, and_most_sig_byte_decomp_0_to_6, and_most_sig_byte_decomp_0_to_7;
};
template <class T> struct IsZeroOperation { T inverse; T result; };
template <class T> struct IsZeroWordOperation {
  IsZeroOperation<T> is_zero_byte[WORD_SIZE];
  T is_lower_half_zero, is_upper_half_zero, result;
};
template <class T> struct IsEqualWordOperation { IsZeroWordOperation<T> is_diff_zero; };
template <class T> struct AddDoubleOperation { Word<T> value; Word<T> value_hi; T carry[7]; };
template <class T> struct AddOperation { Word<T> value; T carry[3]; };   // operations/add.rs:13-19

// crates/core/machine/src/alu/add_sub/mod.rs:43-65
template <class T> struct AddSubCols {
  T pc, next_pc;
  AddOperation<T> add_operation;
  Word<T> operand_1, operand_2;
  T is_add, is_sub;
};
// crates/core/machine/src/alu/bitwise/mod.rs (BitwiseCols)
template <class T> struct BitwiseCols {
  T pc, next_pc;
  Word<T> a, b, c;
  T is_nor, is_xor, is_or, is_and;
};
// crates/core/machine/src/alu/lt/mod.rs (LtCols)
template <class T> struct LtCols {
  T pc, next_pc, is_slt, is_sltu;
  Word<T> a, b, c;
  T byte_flags[4];
  T b_masked, c_masked, not_eq_inv;
  T msb_b, msb_c, bit_b, bit_c;
  T sltu, is_comp_eq, is_sign_eq;
  T comparison_bytes[2];
};
// crates/core/machine/src/alu/sll/mod.rs (ShiftLeftCols)
template <class T> struct ShiftLeftCols {
  T pc, next_pc;
  Word<T> a, b, c;
  T c_least_sig_byte[BYTE_SIZE];
  T shift_by_n_bits[BYTE_SIZE];
  T bit_shift_multiplier;
  T bit_shift_result[WORD_SIZE];
  T bit_shift_result_carry[WORD_SIZE];
  T shift_by_n_bytes[WORD_SIZE];
  T is_real;
};
// crates/core/machine/src/alu/sr/mod.rs (ShiftRightCols)
template <class T> struct ShiftRightCols {
  T pc, next_pc;
  Word<T> b, c;
  T shift_by_n_bits[BYTE_SIZE];
  T shift_by_n_bytes[WORD_SIZE];
  T byte_shift_result[LONG_WORD_SIZE];
  T bit_shift_result[LONG_WORD_SIZE];
  T shr_carry_output_carry[LONG_WORD_SIZE];
  T shr_carry_output_shifted_byte[LONG_WORD_SIZE];
  T b_msb;
  T c_least_sig_byte[BYTE_SIZE];
  T is_srl, is_ror, is_sra;
  T is_real;
};
// crates/core/machine/src/alu/clo_clz/mod.rs (CloClzCols)
template <class T> struct CloClzCols {
  T pc, next_pc;
  Word<T> a, b, bb;
  T is_bb_zero, is_clz, is_real;
};

// crates/core/machine/src/control_flow/branch/columns.rs (BranchColumns)
template <class T> struct BranchColumns {
  T pc;
  Word<T> next_pc;
  KoalaBearWordRangeChecker<T> next_pc_range_checker;
  Word<T> target_pc;
  Word<T> next_next_pc;
  KoalaBearWordRangeChecker<T> next_next_pc_range_checker;
  Word<T> op_a_value, op_b_value, op_c_value;
  T is_beq, is_bne, is_bltz, is_blez, is_bgtz, is_bgez;
  T is_branching, a_gt_b, a_lt_b;
};
// crates/core/machine/src/control_flow/jump/columns.rs (JumpColumns)
template <class T> struct JumpColumns {
  T pc;
  Word<T> next_pc;
  KoalaBearWordRangeChecker<T> next_pc_range_checker;
  Word<T> next_next_pc;
  KoalaBearWordRangeChecker<T> next_next_pc_range_checker;
  Word<T> op_a_value, op_b_value, op_c_value;
  T is_jump, is_jumpi, is_jumpdirect;
  KoalaBearWordRangeChecker<T> op_a_range_checker;
};

// crates/core/machine/src/misc/mov_cond/mod.rs:36-65 (MovCondCols)
template <class T> struct MovCondCols {
  T pc, next_pc;
  Word<T> op_a_value, prev_a_value, op_b_value, op_c_value;
  IsZeroWordOperation<T> c_eq_0;
  T is_mne, is_meq, is_wsbh;
};

// ---- memory records and columns named by crates/core/machine/include/memory.hpp ----
// crates/core/executor/src/events/memory.rs:10-19, :46-60, :66-82 (#[repr(C)])
struct MemoryRecord { uint32_t shard, timestamp, value; };
struct MemoryReadRecord { uint32_t value, shard, timestamp, prev_shard, prev_timestamp; };
struct MemoryWriteRecord { uint32_t value, shard, timestamp, prev_value, prev_shard, prev_timestamp; };
// crates/core/executor/src/events/memory.rs:88-97 (MemoryRecordEnum) and its Option form as cbindgen lays tagged unions out
struct MemoryRecordEnum {
  enum class Tag : uint32_t { Read, Write };
  struct Read_Body { MemoryReadRecord _0; };
  struct Write_Body { MemoryWriteRecord _0; };
  Tag tag;
  union { Read_Body read; Write_Body write; };
};
enum class OptionMemoryRecordEnumTag : uint32_t { Read, Write, None };
struct OptionMemoryRecordEnum { OptionMemoryRecordEnumTag tag; MemoryReadRecord read; MemoryWriteRecord write; };
// crates/core/machine/src/memory/consistency/columns.rs:4-51
template <class T> struct MemoryAccessCols { Word<T> value; T prev_shard, prev_clk, compare_clk, diff_16bit_limb, diff_8bit_limb; };
template <class T> struct MemoryReadCols { MemoryAccessCols<T> access; };
template <class T> struct MemoryWriteCols { Word<T> prev_value; MemoryAccessCols<T> access; };
template <class T> struct MemoryReadWriteCols { Word<T> prev_value; MemoryAccessCols<T> access; };

// ---- Mul (crates/core/machine/include/mul.hpp) ----
// crates/core/executor/src/events/instr.rs:47-73 (#[repr(C)])
struct CompAluEvent {
  uint32_t shard, clk, pc, next_pc;
  Opcode opcode;
  uint32_t hi, a, b, c;
  MemoryWriteRecord hi_record;
  bool hi_record_is_real;
};
// crates/core/machine/src/alu/mul/mod.rs:76-139 (MulCols)
template <class T> struct MulCols {
  T pc, next_pc;
  Word<T> hi, a, b, c;
  T carry[PRODUCT_SIZE];
  T product[PRODUCT_SIZE];
  T b_msb, c_msb, b_sign_extend, c_sign_extend;
  T is_mul, is_mult, is_multu, is_real;
  MemoryReadWriteCols<T> op_hi_access;
  T hi_record_is_real;
  T shard, clk;
};

// ---- MemoryInstrs (crates/core/machine/include/memory_instrs.hpp) ----
// crates/core/executor/src/events/instr.rs:114-136 (#[repr(C)])
struct MemInstrEvent {
  uint32_t shard, clk, pc, next_pc;
  Opcode opcode;
  uint32_t a, b, c;
  MemoryRecordEnum mem_access;
  uint32_t prev_a_val;
};
// crates/core/machine/src/memory/instructions/columns.rs:15-117 (MemoryInstructionsColumns)
template <class T> struct MemoryInstructionsColumns {
  T pc, next_pc, shard, clk;
  Word<T> op_a_value, op_b_value, op_c_value;
  T is_lb, is_lbu, is_lh, is_lhu, is_lw, is_lwl, is_lwr, is_ll, is_sb, is_sh, is_sw, is_swl, is_swr, is_sc;
  Word<T> addr_word;
  T addr_aligned, addr_ls_two_bits, ls_bits_is_one, ls_bits_is_two, ls_bits_is_three;
  KoalaBearWordRangeChecker<T> addr_word_range_checker;
  MemoryReadWriteCols<T> memory_access;
  Word<T> prev_a_val;
  Word<T> unsigned_mem_val;
  T most_sig_bit, most_sig_byte, mem_value_is_neg;
  IsZeroOperation<T> most_sig_bytes_zero;
};

// ---- MemoryLocal (crates/core/machine/include/memory_local.hpp) ----
// crates/core/executor/src/events/memory.rs:228-237 (#[repr(C)])
struct MemoryLocalEvent { uint32_t addr; MemoryRecord initial_mem_access; MemoryRecord final_mem_access; };
// crates/core/machine/src/memory/local.rs:29-55 (SingleMemoryLocal)
template <class T> struct SingleMemoryLocal {
  T addr, initial_shard, final_shard, initial_clk, final_clk;
  Word<T> initial_value, final_value;
  T is_real;
};

// ---- Cpu (crates/core/machine/include/cpu.hpp, instruction.hpp) ----
// crates/core/executor/src/lib.rs:40-52 (#[repr(u8)] tag, #[repr(C)] struct)
enum class OptionValTag : uint8_t { Some = 0, None };
struct OptionU32 { OptionValTag tag; uint32_t value; };
// crates/core/executor/src/events/cpu.rs:46-77 (#[repr(C)])
struct CpuEventFfi {
  uint32_t clk, pc, next_pc, next_next_pc;
  uint32_t a; OptionMemoryRecordEnum a_record;
  uint32_t b; OptionMemoryRecordEnum b_record;
  uint32_t c; OptionMemoryRecordEnum c_record;
  OptionU32 hi;
  OptionMemoryRecordEnum hi_record, memory_record;
  uint32_t exit_code;
};
// crates/core/executor/src/instruction.rs:29-46 (#[repr(C)])
struct InstructionFfi { Opcode opcode; uint8_t op_a; uint32_t op_b, op_c; bool imm_b, imm_c; OptionU32 raw; };
// crates/core/machine/src/cpu/columns/instruction.rs:12-28, cpu/columns/mod.rs:18-84
template <class T> struct InstructionCols { T opcode, op_a; Word<T> op_b, op_c; T op_a_0, imm_b, imm_c; };
template <class T> struct CpuCols {
  T shard, clk_16bit_limb, clk_8bit_limb, shard_to_send, clk_to_send, pc, next_pc, next_next_pc;
  InstructionCols<T> instruction;
  T num_extra_cycles, is_rw_a, is_check_memory, is_halt, is_sequential;
  Word<T> op_a_value, hi_or_prev_a;
  MemoryReadWriteCols<T> op_a_access;
  MemoryReadCols<T> op_b_access, op_c_access;
  T is_real, op_a_immutable;
};

// ---- MiscInstrs (crates/core/machine/include/misc_instrs.hpp) ----
// crates/core/executor/src/events/instr.rs:241-261 (#[repr(C)])
struct MiscEvent {
  uint32_t shard, clk, pc, next_pc;
  Opcode opcode;
  uint32_t a, b, c, prev_a;
  MemoryWriteRecord hi_record;
};
// crates/core/machine/src/misc/others/columns/{maddsub,sext,ext,ins,misc_specific,mod}.rs
template <class T> struct MaddsubCols {
  Word<T> mul_lo, mul_hi;
  AddDoubleOperation<T> add_operation;
  Word<T> src2_hi, src2_lo;
  MemoryReadWriteCols<T> op_hi_access;
};
template <class T> struct SextCols { T most_sig_bit, sig_byte; IsEqualWordOperation<T> a_eq_b; T is_seb, is_seh; };
template <class T> struct ExtCols { T lsb, msbd; Word<T> sll_val; };
template <class T> struct InsCols { T lsb, msb; Word<T> ror_val, srl1_val, srl_val, sll_val, add_val; };
template <class T> union MiscSpecificCols { MaddsubCols<T> maddsub; SextCols<T> sext; ExtCols<T> ext; InsCols<T> ins; };
template <class T> struct MiscInstrColumns {
  T shard, clk, pc, next_pc;
  Word<T> op_a_value, prev_a_value, op_b_value, op_c_value;
  MiscSpecificCols<T> misc_specific_columns;
  T is_sext, is_ins, is_ext, is_maddu, is_msubu, is_madd, is_msub, is_teq;
};

// ---- DivRem (crates/core/machine/include/div_rem.hpp) ----
// crates/core/machine/src/alu/divrem/mod.rs:109-204
template <class T> struct DivRemCols {
  T pc, next_pc;
  Word<T> b, c, quotient, remainder, abs_remainder, abs_c, max_abs_c_or_1;
  T c_times_quotient[LONG_WORD_SIZE];
  T carry[LONG_WORD_SIZE];
  IsZeroWordOperation<T> is_c_0;
  T is_div, is_divu, is_mod, is_modu, is_overflow;
  IsEqualWordOperation<T> is_overflow_b, is_overflow_c;
  T b_msb, rem_msb, c_msb, b_neg, rem_neg, c_neg, remainder_check_multiplicity;
  MemoryReadWriteCols<T> op_hi_access;
  T shard, clk;
};

// ---- SyscallCore / SyscallPrecompile / SyscallInstrs (crates/core/machine/include/syscall.hpp, syscall_instrs.hpp) ----
// crates/core/executor/src/events/syscall.rs:8-29 (#[repr(C)])
struct SyscallEvent {
  uint32_t pc, next_pc, shard, clk;
  MemoryWriteRecord a_record;
  bool a_record_is_real;
  uint32_t syscall_id, arg1, arg2;
};
// crates/core/machine/src/syscall/chip.rs:71-107
template <class T> struct SyscallCols { T shard, clk, syscall_id, arg1_lo, arg1_hi, arg2_lo, arg2_hi, result_lo, result_hi, is_linux, is_real; };
// crates/core/machine/src/syscall/instructions/columns.rs:11-59
template <class T> struct SyscallInstrColumns {
  T pc, next_pc, shard, clk, num_extra_cycles, is_halt, is_sys_linux;
  IsZeroOperation<T> is_prev_a1_zero;
  T syscall_id;
  Word<T> op_a_value, op_b_value, op_c_value, prev_a_value;
  IsZeroOperation<T> is_enter_unconstrained, is_hint_len, is_halt_check, is_exit_group_check, is_commit, is_commit_deferred_proofs;
  T index_bitmap[PV_DIGEST_NUM_WORDS];
  KoalaBearWordRangeChecker<T> op_b_range_check, op_c_range_check;
  T op_b_check, op_c_check, is_real;
};

// ---- MemoryGlobalInit / MemoryGlobalFinalize (crates/core/machine/include/memory_global.hpp) ----
// crates/core/executor/src/events/memory.rs:138-149 (#[repr(C)])
struct MemoryInitializeFinalizeEvent { uint32_t addr, value, shard, timestamp; };
// crates/core/machine/src/operations/cmp.rs:295-298, koala_bear_range.rs:10-31
template <class T, size_t N> struct AssertLtColsBits { T bit_flags[N]; };
template <class T> struct KoalaBearBitDecomposition {
  T bits[32];
  T and_most_sig_byte_decomp_0_to_2, and_most_sig_byte_decomp_0_to_3, and_most_sig_byte_decomp_0_to_4,
      and_most_sig_byte_decomp_0_to_5, and_most_sig_byte_decomp_0_to_6, and_most_sig_byte_decomp_0_to_7;
};
// crates/core/machine/src/memory/global.rs:210-245
template <class T> struct MemoryInitCols {
  T shard, timestamp, addr;
  AssertLtColsBits<T, 32> lt_cols;
  KoalaBearBitDecomposition<T> addr_bits;
  T value[32];
  T is_real, is_next_comp;
  IsZeroOperation<T> is_prev_addr_zero;
  T is_first_comp, is_last_addr;
};

}  // namespace zkm_core_machine_sys

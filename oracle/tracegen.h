// CPU restatement of the core ALU chips' trace generation — TEST
// INFRASTRUCTURE, the checker for ziren_b200/csrc/tracegen.cu, never the thing shipped or measured.
// Plain per-byte loops in canonical form (the CUDA path uses branch-free word arithmetic on
// Montgomery residues), each function following the reference's Rust generate_trace / event_to_row:
//   AddSub     crates/core/machine/src/alu/add_sub/mod.rs:94-123, :152-171; operations/add.rs:23-47
//   Bitwise    crates/core/machine/src/alu/bitwise/mod.rs:90-119 (rows), event_to_row below it
//   Lt         crates/core/machine/src/alu/lt/mod.rs:120-148 (rows), event_to_row below it
//   ShiftLeft  crates/core/machine/src/alu/sll/mod.rs:130-178
//   ShiftRight crates/core/machine/src/alu/sr/mod.rs:163-193, event_to_row below it
//   CloClz     crates/core/machine/src/alu/clo_clz/mod.rs:100-166
//   Branch     crates/core/machine/src/control_flow/branch/trace.rs:45-137 (columns.rs)
//   Jump       crates/core/machine/src/control_flow/jump/trace.rs:45-113 (columns.rs)
//   MovCond    crates/core/machine/src/misc/mov_cond/mod.rs:91-165; operations/is_zero_word.rs, is_zero.rs
//   range checker  crates/core/machine/src/operations/koala_bear_word.rs:35-49
// Pinned against the reference's own C++ row fillers (crates/core/machine/include/*.hpp compiled
// into oracle/_ref/libzkref_core.so) and the golden rows generated from them
// (tests/golden/alu_rows.json).
#pragma once
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include "kb.h"

namespace zko {

struct AluEvent { u32 pc, next_pc, opcode, hi, a, b, c; };
// the same seven words read as BranchEvent / JumpEvent {pc, next_pc, next_next_pc, opcode, a, b, c}
struct FlowEvent { u32 pc, next_pc, next_next_pc, opcode, a, b, c; };
enum { T_ADDSUB = 0, T_BITWISE, T_LT, T_SLL, T_SR, T_CLOCLZ, T_BRANCH, T_JUMP, T_MOVCOND, T_NCHIPS };
static const int ALU_WIDTHS[T_NCHIPS] = {19, 18, 32, 44, 67, 17, 62, 66, 32};
enum { K_ADD = 0, K_SUB = 1, K_SLL = 9, K_SRL = 10, K_SRA = 11, K_ROR = 12, K_SLT = 13, K_SLTU = 14, K_AND = 15, K_OR = 16,
       K_XOR = 17, K_NOR = 18, K_CLZ = 19, K_CLO = 20, K_BEQ = 21, K_BGEZ = 22, K_BGTZ = 23, K_BLEZ = 24, K_BLTZ = 25, K_BNE = 26,
       K_JUMP = 27, K_JUMPI = 28, K_JUMPDIRECT = 29, K_MEQ = 50, K_MNE = 51, K_WSBH = 52 };

struct RowWriter {
  u32* r;
  int at = 0;
  void put(u32 canonical) { r[at++] = canonical % P; }
  void flag(bool b) { r[at++] = b ? 1 : 0; }
  void bytes(const unsigned char* b, int n) { for (int i = 0; i < n; i++) put(b[i]); }
  void word(u32 v) { unsigned char b[4]; for (int i = 0; i < 4; i++) b[i] = (unsigned char)(v >> (8 * i)); bytes(b, 4); }
  // KoalaBearWordRangeChecker::populate
  void range_checker(u32 v) {
    u32 bit[8];
    for (int i = 0; i < 8; i++) { bit[i] = (v >> (24 + i)) & 1; put(bit[i]); }
    u32 acc = bit[0] * bit[1];
    put(acc);
    for (int i = 2; i <= 6; i++) { acc *= bit[i]; put(acc); }
  }
};

static inline void flow_row(int chip, const FlowEvent& e, u32* row) {
  RowWriter w{row};
  w.put(e.pc);
  if (chip == T_BRANCH) {
    w.word(e.next_pc); w.range_checker(e.next_pc);
    w.word(e.next_pc + e.c);
    w.word(e.next_next_pc); w.range_checker(e.next_next_pc);
    w.word(e.a); w.word(e.b); w.word(e.c);
    w.flag(e.opcode == K_BEQ); w.flag(e.opcode == K_BNE); w.flag(e.opcode == K_BLTZ);
    w.flag(e.opcode == K_BLEZ); w.flag(e.opcode == K_BGTZ); w.flag(e.opcode == K_BGEZ);
    const bool eq = e.a == e.b, lt = (int32_t)e.a < (int32_t)e.b, gt = (int32_t)e.a > (int32_t)e.b;
    bool taken = false;
    switch (e.opcode) {
      case K_BEQ: taken = eq; break;
      case K_BNE: taken = !eq; break;
      case K_BLTZ: taken = lt; break;
      case K_BLEZ: taken = lt || eq; break;
      case K_BGTZ: taken = gt; break;
      case K_BGEZ: taken = eq || gt; break;
      default: break;
    }
    w.flag(taken); w.flag(gt); w.flag(lt);
  } else {
    w.word(e.next_pc); w.range_checker(e.next_pc);
    w.word(e.next_next_pc); w.range_checker(e.next_next_pc);
    w.word(e.a); w.word(e.b); w.word(e.c);
    w.flag(e.opcode == K_JUMP); w.flag(e.opcode == K_JUMPI); w.flag(e.opcode == K_JUMPDIRECT);
    w.range_checker(e.a);
  }
  if (w.at != ALU_WIDTHS[chip]) throw std::runtime_error("oracle: control-flow row width mismatch");
}

static inline void alu_row(int chip, const AluEvent& e, u32* row) {
  RowWriter w{row};
  unsigned char a[4], b[4], c[4];
  for (int i = 0; i < 4; i++) { a[i] = (unsigned char)(e.a >> (8 * i)); b[i] = (unsigned char)(e.b >> (8 * i)); c[i] = (unsigned char)(e.c >> (8 * i)); }
  w.put(e.pc); w.put(e.next_pc);
  switch (chip) {
    case T_ADDSUB: {
      const bool is_add = e.opcode == K_ADD;
      const u32 x = is_add ? e.b : e.a, y = e.c;
      w.word(x + y);
      unsigned carry = 0;
      for (int i = 0; i < 3; i++) {
        carry = (((x >> (8 * i)) & 0xff) + ((y >> (8 * i)) & 0xff) + carry) > 0xff;
        w.flag(carry);
      }
      w.word(x); w.word(y);
      w.flag(is_add); w.flag(e.opcode == K_SUB);
      break;
    }
    case T_BITWISE:
      w.bytes(a, 4); w.bytes(b, 4); w.bytes(c, 4);
      w.flag(e.opcode == K_NOR); w.flag(e.opcode == K_XOR); w.flag(e.opcode == K_OR); w.flag(e.opcode == K_AND);
      break;
    case T_LT: {
      const bool slt = e.opcode == K_SLT;
      w.flag(slt); w.flag(e.opcode == K_SLTU);
      w.bytes(a, 4); w.bytes(b, 4); w.bytes(c, 4);
      unsigned char bc[4], cc[4];
      memcpy(bc, b, 4); memcpy(cc, c, 4);
      const unsigned char mb = b[3] & 0x7f, mc = c[3] & 0x7f;
      if (slt) { bc[3] = mb; cc[3] = mc; }
      u32 flags[4] = {0, 0, 0, 0}, cmp[2] = {0, 0}, inv = 0;
      bool sltu = false, eq = true;
      for (int i = 3; i >= 0; i--) {
        if (bc[i] != cc[i]) {
          flags[i] = 1; eq = false;
          sltu = bc[i] < cc[i];
          inv = finv(F(bc[i]) - F(cc[i])).v;
          cmp[0] = bc[i]; cmp[1] = cc[i];
          break;
        }
      }
      for (int i = 0; i < 4; i++) w.put(flags[i]);
      w.put(mb); w.put(mc); w.put(inv);
      const u32 msb_b = b[3] >> 7, msb_c = c[3] >> 7;
      w.put(msb_b); w.put(msb_c); w.put(msb_b * (slt ? 1 : 0)); w.put(msb_c * (slt ? 1 : 0));
      w.flag(sltu); w.flag(eq); w.flag(slt ? msb_b == msb_c : true);
      w.put(cmp[0]); w.put(cmp[1]);
      break;
    }
    case T_SLL: {
      w.bytes(a, 4); w.bytes(b, 4); w.bytes(c, 4);
      for (int i = 0; i < 8; i++) w.flag((e.c >> i) & 1);
      const u32 nbits = e.c % 8, nbytes = (e.c & 31) / 8, mult = 1u << nbits;
      for (u32 i = 0; i < 8; i++) w.flag(nbits == i);
      w.put(mult);
      u32 carry = 0, res[4], car[4];
      for (int i = 0; i < 4; i++) { u32 v = b[i] * mult + carry; carry = v / 256; res[i] = v % 256; car[i] = carry; }
      for (int i = 0; i < 4; i++) w.put(res[i]);
      for (int i = 0; i < 4; i++) w.put(car[i]);
      for (u32 i = 0; i < 4; i++) w.flag(nbytes == i);
      w.flag(true);
      break;
    }
    case T_SR: {
      w.bytes(b, 4); w.bytes(c, 4);
      const u32 n = e.c % 32, nbytes = n / 8, nbits = n % 8;
      for (u32 i = 0; i < 8; i++) w.flag(nbits == i);
      for (u32 i = 0; i < 4; i++) w.flag(nbytes == i);
      unsigned char ext[8];
      for (int i = 0; i < 4; i++) ext[i] = b[i];
      for (int i = 4; i < 8; i++) ext[i] = e.opcode == K_SRA ? ((b[3] >> 7) ? 0xff : 0) : e.opcode == K_ROR ? b[i - 4] : 0;
      unsigned char by[8] = {0};
      for (u32 i = 0; i + nbytes < 8; i++) by[i] = ext[i + nbytes];
      unsigned char out[8], cars[8], shifted[8];
      u32 last = 0;
      for (int i = 7; i >= 0; i--) {
        const unsigned char sh = nbits ? (unsigned char)(by[i] >> nbits) : by[i];
        const unsigned char ca = nbits ? (unsigned char)(by[i] & ((1u << nbits) - 1)) : 0;
        cars[i] = ca; shifted[i] = sh;
        out[i] = (unsigned char)((sh + last * (1u << (8 - nbits))) & 0xff);
        last = ca;
      }
      w.bytes(by, 8); w.bytes(out, 8); w.bytes(cars, 8); w.bytes(shifted, 8);
      w.put(b[3] >> 7);
      for (int i = 0; i < 8; i++) w.flag((e.c >> i) & 1);
      w.flag(e.opcode == K_SRL); w.flag(e.opcode == K_ROR); w.flag(e.opcode == K_SRA);
      w.flag(true);
      break;
    }
    case T_CLOCLZ: {
      const bool clz = e.opcode == K_CLZ;
      const u32 bb = clz ? e.b : 0xffffffffu - e.b;
      w.bytes(a, 4); w.bytes(b, 4); w.word(bb);
      w.flag(bb == 0); w.flag(clz); w.flag(true);
      break;
    }
    default: throw std::runtime_error("oracle: unknown ALU chip");
  }
  if (w.at != ALU_WIDTHS[chip]) throw std::runtime_error("oracle: ALU row width mismatch");
}

// the seven words read as MovCondEvent {pc, next_pc, opcode, a, b, c, prev_a}
static inline void mov_cond_row(const u32* e, u32* row) {
  RowWriter w{row};
  const u32 opcode = e[2] & 0xff, a = e[3], b = e[4], c = e[5], prev_a = e[6];
  w.put(e[0]); w.put(e[1]);
  w.word(a); w.word(prev_a); w.word(b); w.word(c);
  bool zero[4];
  for (int i = 0; i < 4; i++) {          // IsZeroOperation per byte: inverse (0 for a zero byte), result
    const u32 byte = (c >> (8 * i)) & 0xff;
    zero[i] = byte == 0;
    w.put(zero[i] ? 0 : finv(F(byte)).v);
    w.flag(zero[i]);
  }
  w.flag(zero[0] && zero[1]); w.flag(zero[2] && zero[3]); w.flag(zero[0] && zero[1] && zero[2] && zero[3]);
  w.flag(opcode == K_MNE); w.flag(opcode == K_MEQ); w.flag(opcode == K_WSBH);
  if (w.at != ALU_WIDTHS[T_MOVCOND]) throw std::runtime_error("oracle: MovCond row width mismatch");
}

static inline void alu_padding_row(int chip, u32* row) {
  const int w = ALU_WIDTHS[chip];
  for (int i = 0; i < w; i++) row[i] = 0;
  if (chip == T_SLL) { row[22] = 1; row[30] = 1; row[39] = 1; }   // shift_by_n_bits[0], multiplier, shift_by_n_bytes[0]
  if (chip == T_SR) { row[10] = 1; row[18] = 1; }                  // shift_by_n_bits[0], shift_by_n_bytes[0]
  if (chip == T_CLOCLZ) { row[2] = 32; row[14] = 1; }              // a = 32, is_bb_zero
}

// height x width row-major canonical trace: events in order, then padding rows
static inline void alu_trace(int chip, const AluEvent* ev, size_t n, size_t height, u32* out) {
  if (chip < 0 || chip >= T_NCHIPS) throw std::runtime_error("oracle: unknown ALU chip");
  if (n > height) throw std::runtime_error("oracle: more events than rows");
  const int w = ALU_WIDTHS[chip];
  for (size_t i = 0; i < height; i++) {
    if (i >= n) alu_padding_row(chip, out + i * w);
    else if (chip == T_MOVCOND) {
      const u32 words[7] = {ev[i].pc, ev[i].next_pc, ev[i].opcode, ev[i].hi, ev[i].a, ev[i].b, ev[i].c};
      mov_cond_row(words, out + i * w);
    } else if (chip == T_BRANCH || chip == T_JUMP)
      flow_row(chip, FlowEvent{ev[i].pc, ev[i].next_pc, ev[i].opcode, ev[i].hi, ev[i].a, ev[i].b, ev[i].c}, out + i * w);
    else alu_row(chip, ev[i], out + i * w);
  }
}

// ---- Mul (crates/core/machine/src/alu/mul/mod.rs:235-336; C++ twin include/mul.hpp) ---------------------------------
// CompAluEvent (crates/core/executor/src/events/instr.rs:47-73) as 16 words: shard, clk, pc, next_pc, opcode, hi, a, b, c,
// hi_record {value, shard, timestamp, prev_value, prev_shard, prev_timestamp}, hi_record_is_real.
// MulCols (58): pc, next_pc, hi[4], a[4], b[4], c[4], carry[8], product[8], b_msb, c_msb, b_sign_extend, c_sign_extend, is_mul,
// is_mult, is_multu, is_real, op_hi_access {prev_value[4], value[4], prev_shard, prev_clk, compare_clk, diff_16bit_limb,
// diff_8bit_limb}, hi_record_is_real, shard, clk.  Padding rows are zero (mod.rs:165-193).
enum { K_MUL = 2, K_MULT = 3, K_MULTU = 4, MUL_WIDTH = 58, COMP_EVENT_WORDS = 16 };
static inline void mul_row(const u32* e, u32* row) {
  RowWriter w{row};
  const u32 shard = e[0], clk = e[1], pc = e[2], next_pc = e[3], opcode = e[4] & 0xff, hi = e[5], a = e[6], b = e[7], c = e[8];
  const bool hi_real = (e[15] & 0xff) != 0;      // a Rust bool: one byte
  w.put(pc); w.put(next_pc);
  w.word(hi); w.word(a); w.word(b); w.word(c);
  // byte vectors, sign-extended to eight bytes for MULT with a negative operand
  std::vector<u32> bv, cv;
  for (int i = 0; i < 4; i++) { bv.push_back((b >> (8 * i)) & 0xff); cv.push_back((c >> (8 * i)) & 0xff); }
  const u32 b_msb = b >> 31, c_msb = c >> 31;
  const bool b_ext = opcode == K_MULT && b_msb, c_ext = opcode == K_MULT && c_msb;
  if (b_ext) bv.resize(8, 0xff);
  if (c_ext) cv.resize(8, 0xff);
  u32 product[8] = {0}, carry[8] = {0};
  for (size_t i = 0; i < bv.size(); i++)
    for (size_t j = 0; j < cv.size(); j++)
      if (i + j < 8) product[i + j] += bv[i] * cv[j];
  for (int i = 0; i < 8; i++) {
    carry[i] = product[i] / 256;
    product[i] %= 256;
    if (i + 1 < 8) product[i + 1] += carry[i];
  }
  for (int i = 0; i < 8; i++) w.put(carry[i]);
  for (int i = 0; i < 8; i++) w.put(product[i]);
  w.put(b_msb); w.put(c_msb); w.flag(b_ext); w.flag(c_ext);
  w.flag(opcode == K_MUL); w.flag(opcode == K_MULT); w.flag(opcode == K_MULTU); w.flag(true);
  if (hi_real) {
    // MemoryReadWriteCols::populate_write (memory/consistency/trace.rs:44-53): prev_value word, then the access columns
    const u32 value = e[9], rshard = e[10], ts = e[11], prev_value = e[12], prev_shard = e[13], prev_ts = e[14];
    w.word(prev_value);
    w.word(value);
    w.put(prev_shard); w.put(prev_ts);
    const bool same = prev_shard == rshard;
    w.flag(same);
    const u32 d = (same ? ts - prev_ts : rshard - prev_shard) - 1u;
    w.put(d & 0xffff); w.put((d >> 16) & 0xff);
  } else for (int i = 0; i < 13; i++) w.put(0);
  w.flag(hi_real);
  w.put(hi_real ? shard : 0); w.put(hi_real ? clk : 0);
  if (w.at != MUL_WIDTH) throw std::runtime_error("oracle: Mul row width mismatch");
}
static inline void mul_trace(const u32* ev, size_t n, size_t height, u32* out) {
  if (n > height) throw std::runtime_error("oracle: more events than rows");
  for (size_t i = 0; i < height; i++) {
    if (i < n) mul_row(ev + COMP_EVENT_WORDS * i, out + i * MUL_WIDTH);
    else for (int k = 0; k < MUL_WIDTH; k++) out[i * MUL_WIDTH + k] = 0;
  }
}

// ---- MemoryInstrs (crates/core/machine/src/memory/instructions/trace.rs:103-263, columns.rs:15-117; C++ twin
// include/memory_instrs.hpp) ------------------------------------------------------------------------------------------
// MemInstrEvent (crates/core/executor/src/events/instr.rs:114-136) as its 16 #[repr(C)] words: shard, clk, pc, next_pc, opcode,
// a, b, c, mem_access tag (0 Read, 1 Write), the record (MemoryReadRecord {value, shard, timestamp, prev_shard,
// prev_timestamp} or MemoryWriteRecord {value, shard, timestamp, prev_value, prev_shard, prev_timestamp}, six words),
// prev_a_val.  Padding rows are zero (trace.rs:57-80).
enum { K_LB = 31, K_LBU, K_LH, K_LHU, K_LW, K_LWL, K_LWR, K_LL, K_SB, K_SH, K_SW, K_SWL, K_SWR, K_SC, MEMINSTR_WIDTH = 79 };
static inline void mem_instr_row(const u32* e, u32* row) {
  RowWriter w{row};
  const u32 shard = e[0], clk = e[1], pc = e[2], next_pc = e[3], opcode = e[4] & 0xff, a = e[5], b = e[6], c = e[7], tag = e[8];
  const u32* rec = e + 9;
  const u32 prev_a_val = e[15];
  w.put(pc); w.put(next_pc); w.put(shard); w.put(clk);
  w.word(a); w.word(b); w.word(c);
  for (u32 k = K_LB; k <= K_SC; k++) w.flag(opcode == k);
  const u32 memory_addr = b + c;                                   // wrapping_add
  const u32 aligned_addr = memory_addr - memory_addr % 4;
  const u32 ls = memory_addr % 4;
  w.word(memory_addr);
  w.put(aligned_addr); w.put(ls);
  w.flag(ls == 1); w.flag(ls == 2); w.flag(ls == 3);
  w.range_checker(memory_addr);
  // MemoryReadWriteCols::populate (memory/consistency/trace.rs:31-42): current and previous record of either kind
  u32 cur_value, cur_shard, cur_ts, prev_value, prev_shard, prev_ts;
  if (tag == 0) { cur_value = rec[0]; cur_shard = rec[1]; cur_ts = rec[2]; prev_value = rec[0]; prev_shard = rec[3]; prev_ts = rec[4]; }
  else if (tag == 1) { cur_value = rec[0]; cur_shard = rec[1]; cur_ts = rec[2]; prev_value = rec[3]; prev_shard = rec[4]; prev_ts = rec[5]; }
  else throw std::runtime_error("oracle: MemInstrEvent with a bad memory record tag");
  w.word(prev_value);
  w.word(cur_value);
  w.put(prev_shard); w.put(prev_ts);
  const bool use_clk = prev_shard == cur_shard;
  w.flag(use_clk);
  const u32 prev_time = use_clk ? prev_ts : prev_shard, cur_time = use_clk ? cur_ts : cur_shard;
  const u32 diff_minus_one = cur_time - prev_time - 1;
  w.put(diff_minus_one & 0xffff); w.put((diff_minus_one >> 16) & 0xff);
  w.word(prev_a_val);
  // the loaded value before sign extension
  const u32 mem_value = cur_value;
  unsigned char mb[4];
  for (int i = 0; i < 4; i++) mb[i] = (unsigned char)(mem_value >> (8 * i));
  u32 unsigned_mem_val = 0, most_sig_byte = 0, most_sig_bit = 0;
  bool is_neg = false;
  if (opcode >= K_LB && opcode <= K_LL) {
    switch (opcode) {
      case K_LB: case K_LBU: unsigned_mem_val = mb[ls]; break;
      case K_LH: case K_LHU: unsigned_mem_val = ((ls >> 1) % 2 == 0) ? (mem_value & 0x0000FFFF) : ((mem_value & 0xFFFF0000) >> 16); break;
      case K_LW: case K_LL: unsigned_mem_val = mem_value; break;
      case K_LWL: {
        const u32 val = mem_value << (24 - ls * 8), mask = 0xFFFFFFFFu << (24 - ls * 8);
        unsigned_mem_val = (prev_a_val & ~mask) | val;
        break;
      }
      case K_LWR: {
        const u32 val = mem_value >> (ls * 8), mask = 0xFFFFFFFFu >> (ls * 8);
        unsigned_mem_val = (prev_a_val & ~mask) | val;
        break;
      }
    }
    if (opcode == K_LB || opcode == K_LH) {
      most_sig_byte = opcode == K_LB ? (unsigned_mem_val & 0xff) : ((unsigned_mem_val >> 8) & 0xff);
      most_sig_bit = most_sig_byte >> 7;
      is_neg = most_sig_bit == 1;
    }
  }
  w.word(unsigned_mem_val);
  w.put(most_sig_bit); w.put(most_sig_byte); w.flag(is_neg);
  // IsZeroOperation::populate_from_field_element on addr_word[1] + addr_word[2] + addr_word[3] (operations/is_zero.rs)
  const u32 upper = ((memory_addr >> 8) & 0xff) + ((memory_addr >> 16) & 0xff) + ((memory_addr >> 24) & 0xff);
  w.put(upper ? finv(F(upper)).v : 0);
  w.flag(upper == 0);
  if (w.at != MEMINSTR_WIDTH) throw std::runtime_error("oracle: MemoryInstrs row width mismatch");
}
static inline void mem_instr_trace(const u32* ev, size_t n, size_t height, u32* out) {
  if (n > height) throw std::runtime_error("oracle: more events than rows");
  for (size_t i = 0; i < height; i++) {
    if (i < n) mem_instr_row(ev + COMP_EVENT_WORDS * i, out + i * MEMINSTR_WIDTH);
    else for (int k = 0; k < MEMINSTR_WIDTH; k++) out[i * MEMINSTR_WIDTH + k] = 0;
  }
}

// ---- MemoryLocal (crates/core/machine/src/memory/local.rs:146-190; C++ twin of one entry include/memory_local.hpp) -----
// MemoryLocalEvent (crates/core/executor/src/events/memory.rs:228-237) as seven words: addr, initial_mem_access {shard,
// timestamp, value}, final_mem_access {shard, timestamp, value}.  Four events per row (NUM_LOCAL_MEMORY_ENTRIES_PER_ROW);
// SingleMemoryLocal (local.rs:29-55): addr, initial_shard, final_shard, initial_clk, final_clk, initial_value[4],
// final_value[4], is_real.  Entries and rows past the last event are zero.
enum { MEMLOCAL_ENTRIES = 4, MEMLOCAL_ENTRY_WIDTH = 14, MEMLOCAL_WIDTH = 56, MEMLOCAL_EVENT_WORDS = 7 };
static inline void memory_local_entry(const u32* e, u32* cols) {
  RowWriter w{cols};
  const u32 addr = e[0], i_shard = e[1], i_ts = e[2], i_value = e[3], f_shard = e[4], f_ts = e[5], f_value = e[6];
  w.put(addr);
  w.put(i_shard); w.put(f_shard);
  w.put(i_ts); w.put(f_ts);
  w.word(i_value); w.word(f_value);
  w.flag(true);
  if (w.at != MEMLOCAL_ENTRY_WIDTH) throw std::runtime_error("oracle: MemoryLocal entry width mismatch");
}
static inline void memory_local_trace(const u32* ev, size_t n, size_t height, u32* out) {
  const size_t nb_rows = (n + MEMLOCAL_ENTRIES - 1) / MEMLOCAL_ENTRIES;
  if (nb_rows > height) throw std::runtime_error("oracle: more events than rows");
  for (size_t i = 0; i < height * MEMLOCAL_WIDTH; i++) out[i] = 0;
  for (size_t row = 0; row < nb_rows; row++)
    for (size_t k = 0; k < MEMLOCAL_ENTRIES; k++) {
      const size_t idx = row * MEMLOCAL_ENTRIES + k;
      if (idx < n) memory_local_entry(ev + MEMLOCAL_EVENT_WORDS * idx, out + row * MEMLOCAL_WIDTH + k * MEMLOCAL_ENTRY_WIDTH);
    }
}

// ---- Cpu (crates/core/machine/src/cpu/trace.rs:45-246, cpu/columns/mod.rs:18-84, columns/instruction.rs:12-44; C++ twin
// include/cpu.hpp) -------------------------------------------------------------------------------------------------------
// One 28-word record per CpuEvent + its Instruction (include/zkb200.h zkb200_cpu_event): clk, pc, next_pc, next_next_pc, a, b,
// c, hi, flags, opcode | op_a << 8 | shard << 16, op_b, op_c, a_record[6], b_record[5], c_record[5].  Padding rows:
// imm_b = imm_c = is_rw_a = 1 (trace.rs:60-63).
enum { CPU_WIDTH = 67, CPU_EVENT_WORDS = 28, K_SYSCALL = 30, K_INS = 45, K_MADDU = 46, K_MSUBU = 47, K_MADD = 48, K_MSUB = 49,
       K_TEQ = 54, K_DIV = 5, K_DIVU = 6 };
struct CpuInstr {
  u32 opcode;
  bool is_memory_load() const { return opcode >= K_LB && opcode <= K_LL; }
  bool is_memory_store() const { return opcode >= K_SB && opcode <= K_SC; }
  bool is_memory_store_except_sc() const { return is_memory_store() && opcode != K_SC; }
  bool is_maddsub() const { return opcode == K_MADDU || opcode == K_MSUBU || opcode == K_MADD || opcode == K_MSUB; }
  bool is_syscall() const { return opcode == K_SYSCALL; }
  bool is_check_memory() const { return is_syscall() || is_maddsub() || is_memory_load() || is_memory_store(); }
  bool is_rw_a() const { return is_check_memory() || opcode == K_INS || opcode == K_MEQ || opcode == K_MNE; }
  bool is_branch() const { return opcode >= K_BEQ && opcode <= K_BNE; }
  bool is_jump() const { return opcode == K_JUMP || opcode == K_JUMPI || opcode == K_JUMPDIRECT; }
  bool is_mult_div() const { return opcode == K_MULT || opcode == K_MULTU || opcode == K_DIV || opcode == K_DIVU; }
};
// MemoryAccessCols::populate_access without the value word (memory/consistency/trace.rs:56-98)
static inline void cpu_access_tail(RowWriter& w, u32 shard, u32 ts, u32 prev_shard, u32 prev_ts) {
  w.put(prev_shard); w.put(prev_ts);
  const bool use_clk = shard == prev_shard;
  w.flag(use_clk);
  const u32 diff_minus_one = (use_clk ? ts - prev_ts : shard - prev_shard) - 1;
  w.put(diff_minus_one & 0xffff); w.put((diff_minus_one >> 16) & 0xff);
}
static inline void cpu_row(const u32* e, u32* row) {
  RowWriter w{row};
  const u32 clk = e[0], pc = e[1], next_pc = e[2], next_next_pc = e[3], a = e[4], b = e[5], c = e[6], hi = e[7], flags = e[8];
  const bool hi_some = flags & 1, b_read = (flags >> 3) & 1, c_read = (flags >> 4) & 1, imm_b = (flags >> 5) & 1, imm_c = (flags >> 6) & 1;
  const u32 a_kind = (flags >> 1) & 3;                             // 0 None, 1 Read, 2 Write
  if (a_kind == 3) throw std::runtime_error("oracle: cpu event with a bad a_record kind");
  const CpuInstr ins{e[9] & 0xff};
  const u32 op_a = (e[9] >> 8) & 0xff, shard = e[9] >> 16, op_b = e[10], op_c = e[11];
  const u32 *ra = e + 12, *rb = e + 18, *rc = e + 23;
  const bool send = ins.is_check_memory() || ins.is_mult_div();
  w.put(shard); w.put(clk & 0xffff); w.put((clk >> 16) & 0xff);
  w.put(send ? shard : 0); w.put(send ? clk : 0);
  w.put(pc); w.put(next_pc); w.put(next_next_pc);
  w.put(ins.opcode); w.put(op_a); w.word(op_b); w.word(op_c);
  w.flag(op_a == 0); w.flag(imm_b); w.flag(imm_c);
  // the syscall columns need register a's previous value, known once the a record is read
  u32 prev_a = 0, a_value = a, a_tail[5] = {0, 0, 0, 0, 0};
  if (a_kind) {
    RowWriter t{a_tail};
    a_value = ra[0];
    if (a_kind == 1) { prev_a = ra[0]; cpu_access_tail(t, ra[1], ra[2], ra[3], ra[4]); }
    else { prev_a = ra[3]; cpu_access_tail(t, ra[1], ra[2], ra[4], ra[5]); }
  }
  bool is_halt = false;
  u32 num_extra_cycles = 0;
  if (ins.is_syscall()) {
    const u32 id0 = prev_a & 0xff, id1 = (prev_a >> 8) & 0xff, sys_exit_group = 4246 & 0x0FFFF;
    is_halt = (id0 == 0 && id1 == 0) || (id0 == (sys_exit_group & 0xff) && id1 == (sys_exit_group >> 8));
    num_extra_cycles = prev_a >> 24;
  }
  w.put(num_extra_cycles);
  w.flag(ins.is_rw_a()); w.flag(send); w.flag(is_halt);
  w.flag(!is_halt && !ins.is_branch() && !ins.is_jump());
  w.word(a);
  w.word(hi_some ? hi : 0);
  w.word(prev_a); w.word(a_value);
  for (int i = 0; i < 5; i++) w.put(a_tail[i]);
  if (b_read) { w.word(rb[0]); cpu_access_tail(w, rb[1], rb[2], rb[3], rb[4]); } else { w.word(b); for (int i = 0; i < 5; i++) w.put(0); }
  if (c_read) { w.word(rc[0]); cpu_access_tail(w, rc[1], rc[2], rc[3], rc[4]); } else { w.word(c); for (int i = 0; i < 5; i++) w.put(0); }
  w.flag(true);
  w.flag(ins.is_memory_store_except_sc() || ins.is_branch() || ins.opcode == K_TEQ);
  if (w.at != CPU_WIDTH) throw std::runtime_error("oracle: Cpu row width mismatch");
}
static inline void cpu_trace(const u32* ev, size_t n, size_t height, u32* out) {
  if (n > height) throw std::runtime_error("oracle: more events than rows");
  for (size_t i = 0; i < height; i++) {
    u32* r = out + i * CPU_WIDTH;
    if (i < n) { cpu_row(ev + CPU_EVENT_WORDS * i, r); continue; }
    for (int k = 0; k < CPU_WIDTH; k++) r[k] = 0;
    r[19] = 1; r[20] = 1; r[22] = 1;
  }
}

// ---- MiscInstrs (crates/core/machine/src/misc/others/trace.rs:91-275, columns/{mod,maddsub,sext,ext,ins}.rs; C++ twin
// include/misc_instrs.hpp) ---------------------------------------------------------------------------------------------
// MiscEvent (crates/core/executor/src/events/instr.rs:241-261) as 15 words: shard, clk, pc, next_pc, opcode, a, b, c, prev_a,
// hi_record {value, shard, timestamp, prev_value, prev_shard, prev_timestamp}.  72 columns; the 44 columns after op_c_value
// are a union read as MaddsubCols / SextCols / ExtCols / InsCols by opcode.  Padding rows are zero.
enum { MISC_WIDTH = 72, MISC_EVENT_WORDS = 15, MISC_UNION = 44, K_EXT = 53, K_SEXT = 55 };
static inline void misc_row(const u32* e, u32* row) {
  const u32 shard = e[0], clk = e[1], pc = e[2], next_pc = e[3], opcode = e[4] & 0xff, a = e[5], b = e[6], c = e[7], prev_a = e[8];
  const u32 hi_value = e[9], hi_shard = e[10], hi_ts = e[11], hi_prev_value = e[12], hi_prev_shard = e[13], hi_prev_ts = e[14];
  RowWriter w{row};
  w.put(shard); w.put(clk); w.put(pc); w.put(next_pc);
  w.word(a); w.word(prev_a); w.word(b); w.word(c);
  u32 un[MISC_UNION] = {0};
  RowWriter u{un};
  if (opcode == K_SEXT || opcode == K_TEQ) {
    // populate_sext: SextCols {most_sig_bit, sig_byte, a_eq_b, is_seb, is_seh}
    const bool seh = c > 0;
    const u32 sig_bit = seh ? ((b & 0xffff) >> 15) : ((b & 0xff) >> 7), sig_byte = seh ? ((b >> 8) & 0xff) : (b & 0xff);
    u.put(sig_bit); u.put(sig_byte);
    // IsEqualWordOperation -> IsZeroWordOperation over the byte differences as field elements (operations/is_equal_word.rs,
    // is_zero_word.rs, is_zero.rs)
    bool zero[4];
    for (int i = 0; i < 4; i++) {
      const F diff = F((a >> (8 * i)) & 0xff) - F((b >> (8 * i)) & 0xff);
      zero[i] = diff.is_zero();
      u.put(zero[i] ? 0 : finv(diff).v);
      u.flag(zero[i]);
    }
    u.flag(zero[0] && zero[1]); u.flag(zero[2] && zero[3]); u.flag(zero[0] && zero[1] && zero[2] && zero[3]);
    u.flag(!seh); u.flag(seh);
  } else if (opcode == K_MADDU || opcode == K_MSUBU || opcode == K_MADD || opcode == K_MSUB) {
    // populate_maddsub: MaddsubCols {mul_lo, mul_hi, add_operation, src2_hi, src2_lo, op_hi_access}
    const bool is_sign = opcode == K_MADD || opcode == K_MSUB, is_add = opcode == K_MADDU || opcode == K_MADD;
    const u64 multiply = is_sign ? (u64)((long long)(int)b * (long long)(int)c) : (u64)b * (u64)c;
    u.word((u32)multiply); u.word((u32)(multiply >> 32));
    const u32 src2_lo = is_add ? prev_a : a, src2_hi = is_add ? hi_prev_value : hi_value;
    const u64 src2 = ((u64)src2_hi << 32) + src2_lo, expected = multiply + src2;
    u.word((u32)expected); u.word((u32)(expected >> 32));
    u32 carry = 0;
    for (int i = 0; i < 7; i++) {                 // AddDoubleOperation::populate (operations/adddouble.rs:24-64)
      carry = (((multiply >> (8 * i)) & 0xff) + ((src2 >> (8 * i)) & 0xff) + carry) > 255 ? 1 : 0;
      u.put(carry);
    }
    u.word(src2_hi); u.word(src2_lo);
    u.word(hi_prev_value); u.word(hi_value);
    u.put(hi_prev_shard); u.put(hi_prev_ts);
    const bool use_clk = hi_shard == hi_prev_shard;
    u.flag(use_clk);
    const u32 diff_minus_one = (use_clk ? hi_ts - hi_prev_ts : hi_shard - hi_prev_shard) - 1;
    u.put(diff_minus_one & 0xffff); u.put((diff_minus_one >> 16) & 0xff);
  } else if (opcode == K_EXT) {
    const u32 lsb = c & 0x1f, msbd = c >> 5;
    if (lsb + msbd > 31) throw std::runtime_error("oracle: EXT event with lsb + msbd > 31");
    u.put(lsb); u.put(msbd); u.word(b << (31 - lsb - msbd));
  } else if (opcode == K_INS) {
    const u32 lsb = c & 0x1f, msb = c >> 5;
    if (msb < lsb || msb > 31) throw std::runtime_error("oracle: INS event with msb < lsb");
    const u32 ror_val = lsb ? ((prev_a >> lsb) | (prev_a << (32 - lsb))) : prev_a;
    const u32 srl1_val = ror_val >> 1, srl_val = srl1_val >> (msb - lsb), sll_val = b << (31 - msb + lsb);
    u.put(lsb); u.put(msb);
    u.word(ror_val); u.word(srl1_val); u.word(srl_val); u.word(sll_val); u.word(srl_val + sll_val);
  }
  for (int i = 0; i < MISC_UNION; i++) w.put(un[i]);
  w.flag(opcode == K_SEXT); w.flag(opcode == K_INS); w.flag(opcode == K_EXT); w.flag(opcode == K_MADDU);
  w.flag(opcode == K_MSUBU); w.flag(opcode == K_MADD); w.flag(opcode == K_MSUB); w.flag(opcode == K_TEQ);
  if (w.at != MISC_WIDTH) throw std::runtime_error("oracle: MiscInstrs row width mismatch");
}
static inline void misc_trace(const u32* ev, size_t n, size_t height, u32* out) {
  if (n > height) throw std::runtime_error("oracle: more events than rows");
  for (size_t i = 0; i < height; i++) {
    if (i < n) misc_row(ev + MISC_EVENT_WORDS * i, out + i * MISC_WIDTH);
    else for (int k = 0; k < MISC_WIDTH; k++) out[i * MISC_WIDTH + k] = 0;
  }
}

// ---- shared column groups of the chips below ------------------------------------------------------------------------------
// IsZeroOperation::populate_from_field_element (operations/is_zero.rs:29-40): inverse, result
static inline void put_is_zero(RowWriter& w, F a) {
  w.put(a.is_zero() ? 0 : finv(a).v);
  w.flag(a.is_zero());
}
// IsZeroWordOperation::populate_from_field_element (operations/is_zero_word.rs): four IsZeroOperations over the bytes,
// is_lower_half_zero, is_upper_half_zero, result
static inline void put_is_zero_word(RowWriter& w, const F bytes[4]) {
  bool z[4];
  for (int i = 0; i < 4; i++) { z[i] = bytes[i].is_zero(); put_is_zero(w, bytes[i]); }
  w.flag(z[0] && z[1]); w.flag(z[2] && z[3]); w.flag(z[0] && z[1] && z[2] && z[3]);
}
// IsEqualWordOperation::populate (operations/is_equal_word.rs): the zero test of the byte-wise differences a - b
static inline void put_is_equal_word(RowWriter& w, u32 a, u32 b) {
  F d[4];
  for (int i = 0; i < 4; i++) d[i] = F((a >> (8 * i)) & 0xff) - F((b >> (8 * i)) & 0xff);
  put_is_zero_word(w, d);
}
// MemoryReadWriteCols::populate of a write record {value, shard, timestamp, prev_value, prev_shard, prev_timestamp}
// (memory/consistency/trace.rs:44-53, :70-100)
static inline void put_write_access(RowWriter& w, const u32* rec) {
  w.word(rec[3]); w.word(rec[0]);
  w.put(rec[4]); w.put(rec[5]);
  const bool use_clk = rec[4] == rec[1];
  w.flag(use_clk);
  const u32 diff_minus_one = (use_clk ? rec[2] - rec[5] : rec[1] - rec[4]) - 1;
  w.put(diff_minus_one & 0xffff); w.put((diff_minus_one >> 16) & 0xff);
}

// ---- DivRem (crates/core/machine/src/alu/divrem/mod.rs:109-204 columns, :229-364 generate_trace; C++ twin include/div_rem.hpp,
// which returns INT32_MAX instead of u32::MAX for c = 0, writes max(1, |c|) into abs_c and divides INT_MIN by -1 natively: the
// Rust is the authority, the twin pins the rows where both agree) -----------------------------------------------------------
// CompAluEvent as for Mul (16 words).  106 columns; padding rows are zero.
enum { DIVREM_WIDTH = 106, K_MOD = 7, K_MODU = 8 };
static inline void div_rem_row(const u32* e, u32* row) {
  const u32 shard = e[0], clk = e[1], pc = e[2], next_pc = e[3], opcode = e[4] & 0xff, b = e[7], c = e[8];
  if (opcode != K_DIV && opcode != K_DIVU && opcode != K_MOD && opcode != K_MODU) throw std::runtime_error("oracle: DivRem event with another opcode");
  const bool is_signed = opcode == K_DIV || opcode == K_MOD;
  // get_quotient_and_remainder (crates/core/executor/src/utils.rs:33-43): wrapping_div / wrapping_rem
  u32 quotient, remainder;
  if (c == 0) { quotient = 0xffffffffu; remainder = b; }
  else if (is_signed) {
    const long long sb = (int)b, sc = (int)c;           // 64-bit so that INT_MIN / -1 wraps instead of trapping
    quotient = (u32)(sb / sc); remainder = (u32)(sb % sc);
  } else { quotient = b / c; remainder = b % c; }
  RowWriter w{row};
  w.put(pc); w.put(next_pc);
  w.word(b); w.word(c); w.word(quotient); w.word(remainder);
  const auto unsigned_abs = [](u32 v) { return (v >> 31) ? (u32)(0 - v) : v; };
  if (is_signed) {
    const u32 abs_c = unsigned_abs(c);
    w.word(unsigned_abs(remainder)); w.word(abs_c); w.word(abs_c > 1 ? abs_c : 1);
  } else {
    w.word(remainder); w.word(c); w.word(c > 1 ? c : 1);
  }
  unsigned char ctq[8], rb[8];
  const u64 prod = is_signed ? (u64)((long long)(int)quotient * (long long)(int)c) : (u64)quotient * (u64)c;
  const u64 rem64 = is_signed ? (u64)(long long)(int)remainder : (u64)remainder;
  for (int i = 0; i < 8; i++) { ctq[i] = (unsigned char)(prod >> (8 * i)); rb[i] = (unsigned char)(rem64 >> (8 * i)); }
  w.bytes(ctq, 8);
  u32 carry[8];
  for (int i = 0; i < 8; i++) {
    u32 x = (u32)ctq[i] + (u32)rb[i];
    if (i > 0) x += carry[i - 1];
    carry[i] = x / 256;
    w.put(carry[i]);
  }
  F cb[4];
  for (int i = 0; i < 4; i++) cb[i] = F((c >> (8 * i)) & 0xff);
  put_is_zero_word(w, cb);                                                   // is_c_0
  w.flag(opcode == K_DIV); w.flag(opcode == K_DIVU); w.flag(opcode == K_MOD); w.flag(opcode == K_MODU);
  w.flag(is_signed && b == 0x80000000u && c == 0xffffffffu);                 // is_overflow
  put_is_equal_word(w, b, 0x80000000u); put_is_equal_word(w, c, 0xffffffffu);
  const u32 b_msb = b >> 31, rem_msb = remainder >> 31, c_msb = c >> 31;
  w.put(b_msb); w.put(rem_msb); w.put(c_msb);
  w.put(is_signed ? b_msb : 0); w.put(is_signed ? rem_msb : 0); w.put(is_signed ? c_msb : 0);
  w.flag(c != 0);                                                            // remainder_check_multiplicity = 1 - is_c_0.result
  if (opcode == K_DIV || opcode == K_DIVU) { put_write_access(w, e + 9); w.put(shard); w.put(clk); }
  else for (int i = 0; i < 15; i++) w.put(0);
  if (w.at != DIVREM_WIDTH) throw std::runtime_error("oracle: DivRem row width mismatch");
}

// ---- SyscallCore / SyscallPrecompile (crates/core/machine/src/syscall/chip.rs:71-107 columns, :184-268 generate_trace; C++ twin
// include/syscall.hpp) ---------------------------------------------------------------------------------------------------------
// SyscallEvent (crates/core/executor/src/events/syscall.rs:8-29) as 14 words: pc, next_pc, shard, clk, a_record {value, shard,
// timestamp, prev_value, prev_shard, prev_timestamp}, a_record_is_real, syscall_id, arg1, arg2.  Core: the events whose
// prev_value has byte 2 = 1 or byte 1 != 0 (the caller filters, chip.rs:233-240), is_linux = byte 1 != 0.  Precompile: the
// record carries what row_fn takes from the PrecompileEvent, in the convention of syscall.hpp precompile_event_to_row:
// prev_value = 1 and value = v0 for PrecompileEvent::Linux, prev_value = 0 otherwise.  11 columns; padding rows are zero.
enum { SYSCALL_WIDTH = 11, SYSCALL_EVENT_WORDS = 14 };
static inline void syscall_row(const u32* e, bool precompile, u32* row) {
  const u32 shard = e[2], clk = e[3], value = e[4], prev_value = e[7], syscall_id = e[11], arg1 = e[12], arg2 = e[13];
  unsigned char a1[4], a2[4], rb[4];
  for (int i = 0; i < 4; i++) { a1[i] = (unsigned char)(arg1 >> (8 * i)); a2[i] = (unsigned char)(arg2 >> (8 * i)); rb[i] = (unsigned char)(value >> (8 * i)); }
  const bool is_linux = precompile ? prev_value == 1 : ((prev_value >> 8) & 0xff) != 0;
  RowWriter w{row};
  w.put(shard); w.put(clk); w.put(syscall_id);
  w.put(a1[0] + a1[1] * 256u); w.put(a1[2] + a1[3] * 256u);
  w.put(a2[0] + a2[1] * 256u); w.put(a2[2] + a2[3] * 256u);
  w.put(is_linux ? rb[0] + rb[1] * 256u : 0); w.put(is_linux ? rb[2] + rb[3] * 256u : 0);
  w.flag(is_linux); w.flag(true);
  if (w.at != SYSCALL_WIDTH) throw std::runtime_error("oracle: Syscall row width mismatch");
}

// ---- SyscallInstrs (crates/core/machine/src/syscall/instructions/trace.rs:89-177, columns.rs:11-59; C++ twin
// include/syscall_instrs.hpp) -----------------------------------------------------------------------------------------------
// SyscallEvent records; 77 columns; padding rows are zero.  SyscallCode::syscall_id() = the low 16 bits of the code
// (crates/core/executor/src/syscalls/code.rs:269-271).
enum { SYSINSTR_WIDTH = 77, S_HALT = 0x00, S_ENTER_UNCONSTRAINED = 0x03, S_COMMIT = 0x10, S_COMMIT_DEFERRED_PROOFS = 0x1a,
       S_SYSHINTLEN = 0xf0, S_SYS_EXT_GROUP = 4246 };
static inline void syscall_instr_row(const u32* e, u32* row) {
  const u32 pc = e[0], next_pc = e[1], shard = e[2], clk = e[3], value = e[4], prev_value = e[7], syscall_code = e[11], arg1 = e[12], arg2 = e[13];
  const u32 id = prev_value & 0xffff;
  unsigned char pa[4];
  for (int i = 0; i < 4; i++) pa[i] = (unsigned char)(prev_value >> (8 * i));
  const bool is_halt = id == S_HALT || id == S_SYS_EXT_GROUP, send_to_table = pa[1] != 0 || pa[2] == 1;
  RowWriter w{row};
  w.put(pc); w.put(next_pc); w.put(shard); w.put(clk);
  w.put(pa[3]);                                     // num_extra_cycles = prev_a_value[3]
  w.flag(is_halt); w.flag((prev_value & 0x0ff00) != 0);
  put_is_zero(w, F(pa[1]));                         // is_prev_a1_zero
  w.put(syscall_code);
  w.word(value); w.word(arg1); w.word(arg2); w.word(prev_value);
  const u32 tested[6] = {S_ENTER_UNCONSTRAINED, S_SYSHINTLEN, S_HALT, S_SYS_EXT_GROUP, S_COMMIT, S_COMMIT_DEFERRED_PROOFS};
  for (int i = 0; i < 6; i++) put_is_zero(w, F(id) - F(tested[i]));
  const bool commits = id == S_COMMIT || id == S_COMMIT_DEFERRED_PROOFS;
  if (commits && arg1 >= 8) throw std::runtime_error("oracle: COMMIT with a digest index past PV_DIGEST_NUM_WORDS");
  for (u32 i = 0; i < 8; i++) w.flag(commits && arg1 == i);
  const bool b_check = send_to_table || is_halt, c_check = send_to_table || id == S_COMMIT_DEFERRED_PROOFS;
  if (b_check) w.range_checker(arg1); else for (int i = 0; i < 14; i++) w.put(0);
  if (c_check) w.range_checker(arg2); else for (int i = 0; i < 14; i++) w.put(0);
  w.flag(b_check); w.flag(c_check); w.flag(true);
  if (w.at != SYSINSTR_WIDTH) throw std::runtime_error("oracle: SyscallInstrs row width mismatch");
}

// ---- MemoryGlobalInit / MemoryGlobalFinalize (crates/core/machine/src/memory/global.rs:115-192 generate_trace, :210-245 columns;
// C++ twin include/memory_global.hpp for the columns of the first, per-event loop) ------------------------------------------
// The trace takes the address-SORTED MemoryInitializeFinalizeEvent vector {addr, value, shard, timestamp}
// (events/memory.rs:138-149) and the previous address of the public values (previous_init_addr_bits /
// previous_finalize_addr_bits recombined); row i compares with event i - 1, row 0 with the public previous address when that
// is not zero.  111 columns; padding rows are zero.
enum { MEMGLOBAL_WIDTH = 111, MEMGLOBAL_EVENT_WORDS = 4 };
static inline void memory_global_trace(const u32* ev, size_t n, u32 previous_addr, size_t height, u32* out) {
  if (n > height) throw std::runtime_error("oracle: more events than rows");
  for (size_t i = 0; i < height; i++) {
    u32* row = out + i * MEMGLOBAL_WIDTH;
    if (i >= n) { for (int k = 0; k < MEMGLOBAL_WIDTH; k++) row[k] = 0; continue; }
    const u32 addr = ev[4 * i], value = ev[4 * i + 1], shard = ev[4 * i + 2], timestamp = ev[4 * i + 3];
    if (i > 0 && ev[4 * (i - 1)] >= addr) throw std::runtime_error("oracle: memory events are not sorted by address");
    u32 lt[32] = {0};
    // AssertLtColsBits::populate (operations/cmp.rs:300-319): from the top bit down, flag the first bit where a < b
    const auto populate_lt = [&](u32 a, u32 b) {
      for (int k = 31; k >= 0; k--) {
        const u32 ak = (a >> k) & 1, bk = (b >> k) & 1;
        if (ak > bk) throw std::runtime_error("oracle: previous address is not below the address");
        if (ak < bk) { lt[k] = 1; break; }
      }
    };
    bool is_next_comp = false, is_first_comp = false;
    u32 prev_inverse = 0; bool prev_zero = false;
    if (i == 0) {
      prev_zero = previous_addr == 0;
      prev_inverse = prev_zero ? 0 : finv(F(previous_addr)).v;
      is_first_comp = !prev_zero;
      if (!prev_zero) populate_lt(previous_addr, addr);
    } else {
      is_next_comp = true;
      populate_lt(ev[4 * (i - 1)], addr);
    }
    RowWriter w{row};
    w.put(shard); w.put(timestamp); w.put(addr);
    for (int k = 0; k < 32; k++) w.put(lt[k]);
    // KoalaBearBitDecomposition::populate (operations/koala_bear_range.rs:34-48)
    u32 bit[32];
    for (int k = 0; k < 32; k++) { bit[k] = (addr >> k) & 1; w.put(bit[k]); }
    u32 acc = bit[24] * bit[25];
    w.put(acc);
    for (int k = 26; k <= 30; k++) { acc *= bit[k]; w.put(acc); }
    for (int k = 0; k < 32; k++) w.put((value >> k) & 1);
    w.flag(true); w.flag(is_next_comp);
    w.put(prev_inverse); w.flag(i == 0 && prev_zero);
    w.flag(is_first_comp); w.flag(i == n - 1);
    if (w.at != MEMGLOBAL_WIDTH) throw std::runtime_error("oracle: MemoryGlobal row width mismatch");
  }
}

// one-event-per-row chips by name: rows of `height` x width canonical words, zero padding rows
static inline int chip_trace_width(const std::string& chip) {
  if (chip == "DivRem") return DIVREM_WIDTH;
  if (chip == "SyscallCore" || chip == "SyscallPrecompile") return SYSCALL_WIDTH;
  if (chip == "SyscallInstrs") return SYSINSTR_WIDTH;
  return -1;
}
static inline int chip_event_words(const std::string& chip) {
  if (chip == "DivRem") return COMP_EVENT_WORDS;
  if (chip == "SyscallCore" || chip == "SyscallPrecompile" || chip == "SyscallInstrs") return SYSCALL_EVENT_WORDS;
  return -1;
}
static inline void chip_trace(const std::string& chip, const u32* ev, size_t n, size_t height, u32* out) {
  const int w = chip_trace_width(chip), ew = chip_event_words(chip);
  if (w < 0) throw std::runtime_error("oracle: no row filler for chip " + chip);
  if (n > height) throw std::runtime_error("oracle: more events than rows");
  for (size_t i = 0; i < height; i++) {
    u32* row = out + i * w;
    if (i >= n) { for (int k = 0; k < w; k++) row[k] = 0; continue; }
    const u32* e = ev + (size_t)ew * i;
    if (chip == "DivRem") div_rem_row(e, row);
    else if (chip == "SyscallInstrs") syscall_instr_row(e, row);
    else syscall_row(e, chip == "SyscallPrecompile", row);
  }
}

}  // namespace zko

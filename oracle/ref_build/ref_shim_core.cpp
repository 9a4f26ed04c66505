// Drives the reference's own ALU row fillers (crates/core/machine/include/*.hpp, the C++ the
// reference's `sys` feature calls from generate_trace, crates/core/machine/cpp/extern.cpp:15-80)
// so that tests can compare our trace generation with it.  Test infrastructure only.
#include <cassert>
#include <algorithm>
#include <cstring>
#include <string>
#include "zkm-core-machine-sys-cbindgen.hpp"
#include "kb31_t.hpp"
#include "add_sub.hpp"
#include "bitwise.hpp"
#include "lt.hpp"
#include "shift_left.hpp"
#include "shift_right.hpp"
#include "clo_clz.hpp"
#include "branch.hpp"
#include "jump.hpp"
#include "mov_cond.hpp"
#include "memory.hpp"
#include "mul.hpp"
#include "memory_instrs.hpp"
#include "memory_local.hpp"
#include "cpu.hpp"
#include "misc_instrs.hpp"
#include "div_rem.hpp"
#include "syscall.hpp"
#include "syscall_instrs.hpp"
#include "memory_global.hpp"

using namespace zkm_core_machine_sys;

template <class Cols> static constexpr size_t ncols() { return sizeof(Cols) / sizeof(kb31_t); }

extern "C" {
// chip: 0 AddSub, 1 Bitwise, 2 Lt, 3 ShiftLeft, 4 ShiftRight, 5 CloClz, 6 Branch, 7 Jump, 8 MovCond
unsigned ref_alu_num_cols(int chip) {
  switch (chip) {
    case 0: return ncols<AddSubCols<kb31_t>>();
    case 1: return ncols<BitwiseCols<kb31_t>>();
    case 2: return ncols<LtCols<kb31_t>>();
    case 3: return ncols<ShiftLeftCols<kb31_t>>();
    case 4: return ncols<ShiftRightCols<kb31_t>>();
    case 5: return ncols<CloClzCols<kb31_t>>();
    case 6: return ncols<BranchColumns<kb31_t>>();
    case 7: return ncols<JumpColumns<kb31_t>>();
    case 8: return ncols<MovCondCols<kb31_t>>();
  }
  return 0;
}
// events: n records of 7 words, AluEvent {pc, next_pc, opcode, hi, a, b, c} for chips 0-5,
// BranchEvent / JumpEvent {pc, next_pc, next_next_pc, opcode, a, b, c} for chips 6-7,
// MovCondEvent {pc, next_pc, opcode, a, b, c, prev_a} for chip 8; rows: n x num_cols, zero-initialised here,
// Montgomery words exactly as the reference leaves them in the trace
int ref_alu_event_to_rows(int chip, const uint32_t* ev, size_t n, uint32_t* rows) {
  const unsigned w = ref_alu_num_cols(chip);
  if (!w) return 1;
  std::memset(rows, 0, n * w * sizeof(uint32_t));
  for (size_t i = 0; i < n; i++) {
    AluEvent e;
    e.pc = ev[7 * i]; e.next_pc = ev[7 * i + 1]; e.opcode = (Opcode)ev[7 * i + 2]; e.hi = ev[7 * i + 3];
    e.a = ev[7 * i + 4]; e.b = ev[7 * i + 5]; e.c = ev[7 * i + 6];
    uint32_t* r = rows + i * w;
    BranchEvent be{ev[7 * i], ev[7 * i + 1], ev[7 * i + 2], (Opcode)ev[7 * i + 3], ev[7 * i + 4], ev[7 * i + 5], ev[7 * i + 6]};
    JumpEvent je{be.pc, be.next_pc, be.next_next_pc, be.opcode, be.a, be.b, be.c};
    MovCondEvent me{ev[7 * i], ev[7 * i + 1], (Opcode)ev[7 * i + 2], ev[7 * i + 3], ev[7 * i + 4], ev[7 * i + 5], ev[7 * i + 6]};
    switch (chip) {
      case 0: add_sub::event_to_row<kb31_t>(e, *reinterpret_cast<AddSubCols<kb31_t>*>(r)); break;
      case 1: bitwise::event_to_row<kb31_t>(e, *reinterpret_cast<BitwiseCols<kb31_t>*>(r)); break;
      case 2: lt::event_to_row<kb31_t>(e, *reinterpret_cast<LtCols<kb31_t>*>(r)); break;
      case 3: shift_left::event_to_row<kb31_t>(e, *reinterpret_cast<ShiftLeftCols<kb31_t>*>(r)); break;
      case 4: shift_right::event_to_row<kb31_t>(e, *reinterpret_cast<ShiftRightCols<kb31_t>*>(r)); break;
      case 5: clo_clz::event_to_row<kb31_t>(e, *reinterpret_cast<CloClzCols<kb31_t>*>(r)); break;
      case 6: branch::event_to_row<kb31_t>(be, *reinterpret_cast<BranchColumns<kb31_t>*>(r)); break;
      case 7: jump::event_to_row<kb31_t>(je, *reinterpret_cast<JumpColumns<kb31_t>*>(r)); break;
      case 8: mov_cond::event_to_row<kb31_t>(me, *reinterpret_cast<MovCondCols<kb31_t>*>(r)); break;
    }
  }
  return 0;
}
// MemoryReadCols::populate of the reference's memory.hpp:36-52 (9 Montgomery words: value[4], prev_shard, prev_clk,
// compare_clk, diff_16bit_limb, diff_8bit_limb)
void ref_mem_read_cols(uint32_t value, uint32_t shard, uint32_t ts, uint32_t prev_shard, uint32_t prev_ts, uint32_t* out9) {
  MemoryReadCols<kb31_t> c;
  std::memset(&c, 0, sizeof(c));
  MemoryReadRecord r{value, shard, ts, prev_shard, prev_ts};
  memory::populate_read<kb31_t>(c, r);
  static_assert(sizeof(c) == 9 * sizeof(uint32_t), "MemoryReadCols is nine field elements");
  std::memcpy(out9, &c, sizeof(c));
}
// MemoryReadWriteCols through populate_read_write_v2 with a write record (13 words: prev_value[4], access[9]) =
// what MemoryWriteCols::populate leaves (memory/consistency/trace.rs:8-20)
void ref_mem_write_cols(uint32_t value, uint32_t shard, uint32_t ts, uint32_t prev_value, uint32_t prev_shard, uint32_t prev_ts, uint32_t* out13) {
  MemoryReadWriteCols<kb31_t> c;
  std::memset(&c, 0, sizeof(c));
  MemoryRecordEnum e;
  e.tag = MemoryRecordEnum::Tag::Write;
  e.write._0 = MemoryWriteRecord{value, shard, ts, prev_value, prev_shard, prev_ts};
  memory::populate_read_write_v2<kb31_t>(c, e);
  std::memcpy(out13, &c, sizeof(c));
}
// MulChip rows of the reference's mul.hpp: events n x 16 words {shard, clk, pc, next_pc, opcode, hi, a, b, c,
// hi_record{value, shard, timestamp, prev_value, prev_shard, prev_timestamp}, hi_record_is_real}; rows n x 58 Montgomery words
unsigned ref_mul_num_cols() { return ncols<MulCols<kb31_t>>(); }
int ref_mul_event_to_rows(const uint32_t* ev, size_t n, uint32_t* rows) {
  const unsigned w = ref_mul_num_cols();
  std::memset(rows, 0, n * w * sizeof(uint32_t));
  for (size_t i = 0; i < n; i++) {
    const uint32_t* e = ev + 16 * i;
    CompAluEvent c{e[0], e[1], e[2], e[3], (Opcode)e[4], e[5], e[6], e[7], e[8],
                   MemoryWriteRecord{e[9], e[10], e[11], e[12], e[13], e[14]}, e[15] != 0};
    mul::event_to_row<kb31_t>(c, *reinterpret_cast<MulCols<kb31_t>*>(rows + i * w));
  }
  return 0;
}
// MemoryInstrs rows of the reference's memory_instrs.hpp: events n x 16 words, the #[repr(C)] image of MemInstrEvent {shard,
// clk, pc, next_pc, opcode, a, b, c, mem_access{tag, six record words}, prev_a_val}; rows n x 79 Montgomery words
unsigned ref_mem_instr_num_cols() { return ncols<MemoryInstructionsColumns<kb31_t>>(); }
int ref_mem_instr_event_to_rows(const uint32_t* ev, size_t n, uint32_t* rows) {
  static_assert(sizeof(MemInstrEvent) == 16 * sizeof(uint32_t), "MemInstrEvent is sixteen words");
  const unsigned w = ref_mem_instr_num_cols();
  std::memset(rows, 0, n * w * sizeof(uint32_t));
  for (size_t i = 0; i < n; i++) {
    const uint32_t* e = ev + 16 * i;
    MemInstrEvent m;
    std::memset(&m, 0, sizeof(m));
    m.shard = e[0]; m.clk = e[1]; m.pc = e[2]; m.next_pc = e[3]; m.opcode = (Opcode)e[4]; m.a = e[5]; m.b = e[6]; m.c = e[7];
    if (e[8] == 0) {
      m.mem_access.tag = MemoryRecordEnum::Tag::Read;
      m.mem_access.read._0 = MemoryReadRecord{e[9], e[10], e[11], e[12], e[13]};
    } else {
      m.mem_access.tag = MemoryRecordEnum::Tag::Write;
      m.mem_access.write._0 = MemoryWriteRecord{e[9], e[10], e[11], e[12], e[13], e[14]};
    }
    m.prev_a_val = e[15];
    memory_instrs::event_to_row<kb31_t>(m, *reinterpret_cast<MemoryInstructionsColumns<kb31_t>*>(rows + i * w));
  }
  return 0;
}
// One SingleMemoryLocal entry (14 Montgomery words) per MemoryLocalEvent (seven words: addr, initial {shard, timestamp,
// value}, final {shard, timestamp, value}) through the reference's memory_local.hpp
int ref_memory_local_entries(const uint32_t* ev, size_t n, uint32_t* out) {
  static_assert(sizeof(MemoryLocalEvent) == 7 * sizeof(uint32_t), "MemoryLocalEvent is seven words");
  static_assert(sizeof(SingleMemoryLocal<kb31_t>) == 14 * sizeof(uint32_t), "SingleMemoryLocal is fourteen field elements");
  std::memset(out, 0, n * 14 * sizeof(uint32_t));
  for (size_t i = 0; i < n; i++) {
    const uint32_t* e = ev + 7 * i;
    MemoryLocalEvent m{e[0], MemoryRecord{e[1], e[2], e[3]}, MemoryRecord{e[4], e[5], e[6]}};
    memory_local::event_to_row<kb31_t, kb31_septic_extension_t>(&m, reinterpret_cast<SingleMemoryLocal<kb31_t>*>(out + 14 * i));
  }
  return 0;
}
// Cpu rows of the reference's cpu.hpp: events n x 28 words (include/zkb200.h zkb200_cpu_event: clk, pc, next_pc, next_next_pc,
// a, b, c, hi, flags, opcode | op_a << 8 | shard << 16, op_b, op_c, a_record[6], b_record[5], c_record[5]) unpacked into the
// CpuEventFfi / InstructionFfi / shard arguments of cpu_event_to_row_koalabear (cpp/extern.cpp:6-14); rows n x 67 Montgomery words
static OptionMemoryRecordEnum option_record(unsigned kind, const uint32_t* r) {
  OptionMemoryRecordEnum o;
  std::memset(&o, 0, sizeof(o));
  o.tag = kind == 1 ? OptionMemoryRecordEnumTag::Read : kind == 2 ? OptionMemoryRecordEnumTag::Write : OptionMemoryRecordEnumTag::None;
  if (kind == 1) o.read = MemoryReadRecord{r[0], r[1], r[2], r[3], r[4]};
  if (kind == 2) o.write = MemoryWriteRecord{r[0], r[1], r[2], r[3], r[4], r[5]};
  return o;
}
unsigned ref_cpu_num_cols() { return ncols<CpuCols<kb31_t>>(); }
int ref_cpu_event_to_rows(const uint32_t* ev, size_t n, uint32_t* rows) {
  const unsigned w = ref_cpu_num_cols();
  std::memset(rows, 0, n * w * sizeof(uint32_t));
  for (size_t i = 0; i < n; i++) {
    const uint32_t* e = ev + 28 * i;
    const uint32_t fl = e[8];
    CpuEventFfi c;
    std::memset(&c, 0, sizeof(c));
    c.clk = e[0]; c.pc = e[1]; c.next_pc = e[2]; c.next_next_pc = e[3]; c.a = e[4]; c.b = e[5]; c.c = e[6];
    c.hi = OptionU32{(fl & 1) ? OptionValTag::Some : OptionValTag::None, e[7]};
    c.a_record = option_record((fl >> 1) & 3, e + 12);
    c.b_record = option_record((fl >> 3) & 1, e + 18);
    c.c_record = option_record((fl >> 4) & 1, e + 23);
    c.hi_record = option_record(0, nullptr);
    c.memory_record = option_record(0, nullptr);
    InstructionFfi ins{(Opcode)(e[9] & 0xff), (uint8_t)((e[9] >> 8) & 0xff), e[10], e[11], ((fl >> 5) & 1) != 0, ((fl >> 6) & 1) != 0,
                       OptionU32{OptionValTag::None, 0}};
    cpu::event_to_row<kb31_t>(c, e[9] >> 16, ins, *reinterpret_cast<CpuCols<kb31_t>*>(rows + i * w));
  }
  return 0;
}
// MiscInstrs rows of the reference's misc_instrs.hpp: events n x 15 words, the #[repr(C)] image of MiscEvent {shard, clk, pc,
// next_pc, opcode, a, b, c, prev_a, hi_record[6]}; rows n x 72 Montgomery words
unsigned ref_misc_num_cols() { return ncols<MiscInstrColumns<kb31_t>>(); }
int ref_misc_event_to_rows(const uint32_t* ev, size_t n, uint32_t* rows) {
  static_assert(sizeof(MiscEvent) == 15 * sizeof(uint32_t), "MiscEvent is fifteen words");
  const unsigned w = ref_misc_num_cols();
  std::memset(rows, 0, n * w * sizeof(uint32_t));
  for (size_t i = 0; i < n; i++) {
    const uint32_t* e = ev + 15 * i;
    MiscEvent m{e[0], e[1], e[2], e[3], (Opcode)(e[4] & 0xff), e[5], e[6], e[7], e[8], MemoryWriteRecord{e[9], e[10], e[11], e[12], e[13], e[14]}};
    misc_instrs::event_to_row<kb31_t>(m, *reinterpret_cast<MiscInstrColumns<kb31_t>*>(rows + i * w));
  }
  return 0;
}
// rows of the reference's div_rem.hpp / syscall.hpp / syscall_instrs.hpp / memory_global.hpp event_to_row by chip name
// (MachineAir::name); events as the #[repr(C)] words of CompAluEvent (16), SyscallEvent (14), MemoryInitializeFinalizeEvent (4);
// rows n x num_cols Montgomery words.  memory_global.hpp fills only the columns that depend on the event alone.
static SyscallEvent syscall_event_from_words(const uint32_t* e) {
  return SyscallEvent{e[0], e[1], e[2], e[3], MemoryWriteRecord{e[4], e[5], e[6], e[7], e[8], e[9]}, (e[10] & 0xff) != 0, e[11], e[12], e[13]};
}
unsigned ref_chip_num_cols(const char* chip) {
  const std::string c(chip);
  if (c == "DivRem") return ncols<DivRemCols<kb31_t>>();
  if (c == "SyscallCore" || c == "SyscallPrecompile") return ncols<SyscallCols<kb31_t>>();
  if (c == "SyscallInstrs") return ncols<SyscallInstrColumns<kb31_t>>();
  if (c == "MemoryGlobalInit" || c == "MemoryGlobalFinalize") return ncols<MemoryInitCols<kb31_t>>();
  return 0;
}
int ref_chip_event_to_rows(const char* chip, const uint32_t* ev, size_t n, uint32_t* rows) {
  static_assert(sizeof(SyscallEvent) == 14 * sizeof(uint32_t), "SyscallEvent is fourteen words");
  static_assert(sizeof(CompAluEvent) == 16 * sizeof(uint32_t), "CompAluEvent is sixteen words");
  const std::string c(chip);
  const unsigned w = ref_chip_num_cols(chip);
  if (!w) return 1;
  std::memset(rows, 0, n * w * sizeof(uint32_t));
  for (size_t i = 0; i < n; i++) {
    uint32_t* r = rows + i * w;
    if (c == "DivRem") {
      const uint32_t* e = ev + 16 * i;
      CompAluEvent m{e[0], e[1], e[2], e[3], (Opcode)(e[4] & 0xff), e[5], e[6], e[7], e[8],
                     MemoryWriteRecord{e[9], e[10], e[11], e[12], e[13], e[14]}, (e[15] & 0xff) != 0};
      div_rem::event_to_row<kb31_t>(m, *reinterpret_cast<DivRemCols<kb31_t>*>(r));
    } else if (c == "SyscallCore") {
      syscall::core_event_to_row<kb31_t>(syscall_event_from_words(ev + 14 * i), *reinterpret_cast<SyscallCols<kb31_t>*>(r));
    } else if (c == "SyscallPrecompile") {
      syscall::precompile_event_to_row<kb31_t>(syscall_event_from_words(ev + 14 * i), *reinterpret_cast<SyscallCols<kb31_t>*>(r));
    } else if (c == "SyscallInstrs") {
      syscall_instrs::event_to_row<kb31_t>(syscall_event_from_words(ev + 14 * i), *reinterpret_cast<SyscallInstrColumns<kb31_t>*>(r));
    } else {
      const uint32_t* e = ev + 4 * i;
      MemoryInitializeFinalizeEvent m{e[0], e[1], e[2], e[3]};
      memory_global::event_to_row<kb31_t, kb31_septic_extension_t>(&m, c == "MemoryGlobalFinalize", reinterpret_cast<MemoryInitCols<kb31_t>*>(r));
    }
  }
  return 0;
}
// the reference's septic extension and curve (crates/core/machine/include/kb31_septic_extension_t.hpp), canonical words in and
// out: op 0 a * b, 1 reciprocal, 2 sqrt (returns 1 when a is not a square), 3 frobenius, 4 double_frobenius, 5 curve_formula
// (7 words each); 6 point addition (14 words each, infinity as zeros; its doubling branch differs from the Rust and is not used)
int ref_septic_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  kb31_t av[14], bv[14];
  for (int i = 0; i < 14; i++) {
    av[i] = kb31_t::from_canonical_u32(op == 6 || i < 7 ? a[i] : 0);
    bv[i] = kb31_t::from_canonical_u32(b && (op == 6 || i < 7) ? b[i] : 0);
  }
  if (op == 6) {
    kb31_septic_curve_t p(av), q(bv);
    p += q;
    for (int i = 0; i < 7; i++) { out[i] = p.x.value[i].as_canonical_u32(); out[7 + i] = p.y.value[i].as_canonical_u32(); }
    return 0;
  }
  kb31_septic_extension_t x(av), y(bv), r;
  switch (op) {
    case 0: r = x * y; break;
    case 1: r = x.reciprocal(); break;
    case 2: {
      const kb31_t pow_r = x.pow_r();
      if ((pow_r ^ 1065353216) != kb31_t::one()) return 1;
      r = x.sqrt(pow_r);
      break;
    }
    case 3: r = x.frobenius(); break;
    case 4: r = x.double_frobenius(); break;
    case 5: r = x.curve_formula(); break;
    default: return -1;
  }
  for (int i = 0; i < 7; i++) out[i] = r.value[i].as_canonical_u32();
  return 0;
}
}

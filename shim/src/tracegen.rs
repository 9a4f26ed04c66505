//! Trace generation moved to the device: for the chips libzkb200 has a row filler for,
//! the host hands `zkb200_commit` the record's EVENT VECTORS instead of generated rows (`ZKB200_TRACE_EVENTS`,
//! include/zkb200.h).  The table is written column-major straight into the shard's trace storage on the GPU: none of its
//! bytes cross PCIe and no layout change runs (KeccakSponge: 1.5 KB of record per 24 rows x 3531 columns = 339 KB of rows).
//!
//! * ALU / control-flow chips: `AluEvent`, `BranchEvent`, `JumpEvent`, `MovCondEvent` are `#[repr(C)]` seven-word records,
//!   `CompAluEvent` (Mul) and `MemInstrEvent` (MemoryInstrs) sixteen-word ones, `MiscEvent` (MiscInstrs) fifteen words
//!   (crates/core/executor/src/events/instr.rs);
//!   they cross as they lie in `record.add_sub_events` etc. (`event_vector`).  The byte-lookup multiplicities these
//!   chips' `event_to_row` also emits come from `generate_dependencies`, which the caller still runs on the host.
//! * Global: the 32-byte `GlobalLookupEvent` records as they lie; the curve lift (a square root in F_p^7 per trial) and the
//!   running curve sum - the expensive part of `GlobalChip::generate_trace` - run on the device.
//! * DivRem (`CompAluEvent`) and SyscallInstrs (`SyscallEvent`, fourteen words) cross as they lie too; SyscallCore,
//!   SyscallPrecompile and MemoryGlobalInit / MemoryGlobalFinalize take records built here (`owned_events`): the filtered /
//!   normalised syscall events, and the address-sorted memory events with the neighbour's address folded into each record.
//! * MemoryLocal: seven-word `MemoryLocalEvent` records (crates/core/executor/src/events/memory.rs:228-237), four to a row;
//!   the record keeps them in several vectors (`get_local_mem_events`), so they are gathered into one (28 bytes per event
//!   against 224 bytes per row of four).
//! * Cpu: `CpuEvent` (crates/core/executor/src/events/cpu.rs:15-44) holds Options and its instruction lives in the program,
//!   so each event is flattened into a 28-word `zkb200_cpu_event` (`flatten_cpu_events`), the same information the
//!   reference's own FFI passes as (CpuEventFfi, shard, InstructionFfi) (crates/core/machine/src/sys.rs:24-29).
//! * KeccakSponge: `KeccakSpongeEvent` (crates/core/executor/src/events/precompiles/keccak_sponge.rs:15-40) holds Vecs, so
//!   it is flattened into one `zkb200_keccak_block` per absorbed block (24 rows), mirroring the block loop of
//!   `KeccakSpongeChip::event_to_rows` (crates/core/machine/src/syscall/precompiles/keccak_sponge/trace.rs:101-196).
//! NOT compiled in this repository (no Rust toolchain in the build image).
use zkm_core_executor::events::{KeccakSpongeEvent, MemoryReadRecord, MemoryWriteRecord, PrecompileEvent};
use zkm_core_executor::syscalls::SyscallCode;
use zkm_core_executor::ExecutionRecord;

pub const KECCAK_RATE_U32S: usize = 36; // KECCAK_GENERAL_RATE_U32S
pub const KECCAK_OUTPUT_U32S: usize = 16; // KECCAK_GENERAL_OUTPUT_U32S
pub const KECCAK_ROUNDS: usize = 24; // p3_keccak_air::NUM_ROUNDS
pub const NUM_KECCAK_SPONGE_COLS: usize = 3531;

/// `zkb200_keccak_block` (include/zkb200.h): 384 words.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct KeccakBlock {
    pub shard: u32,
    pub clk: u32,
    pub input_addr: u32,
    pub output_addr: u32,
    pub input_len: u32,
    pub block: u32,
    pub num_blocks: u32,
    pub reserved: u32,
    pub xored_state: [u32; 50],
    pub input: [u32; 36],
    pub input_reads: [[u32; 5]; 36],
    pub input_length_read: [u32; 5],
    pub output_writes: [[u32; 6]; 16],
    pub pad: [u32; 9],
}
const _: () = assert!(core::mem::size_of::<KeccakBlock>() == 384 * 4);

fn read_words(r: &MemoryReadRecord) -> [u32; 5] {
    [r.value, r.shard, r.timestamp, r.prev_shard, r.prev_timestamp]
}
fn write_words(w: &MemoryWriteRecord) -> [u32; 6] {
    [w.value, w.shard, w.timestamp, w.prev_value, w.prev_shard, w.prev_timestamp]
}

/// One record per absorbed block of every KECCAK_SPONGE event of the shard, in event order.
pub fn flatten_keccak_sponge_events(record: &ExecutionRecord) -> Vec<KeccakBlock> {
    let mut out = Vec::new();
    for (_, event) in record.get_precompile_events(SyscallCode::KECCAK_SPONGE) {
        let PrecompileEvent::KeccakSponge(event) = event else { unreachable!() };
        flatten_event(event, &mut out);
    }
    out
}

fn flatten_event(event: &KeccakSpongeEvent, out: &mut Vec<KeccakBlock>) {
    let nb = event.num_blocks();
    for i in 0..nb {
        let mut b = KeccakBlock {
            shard: event.shard,
            clk: event.clk,
            input_addr: event.input_addr,
            output_addr: event.output_addr,
            input_len: event.input.len() as u32,
            block: i as u32,
            num_blocks: nb as u32,
            reserved: 0,
            xored_state: [0; 50],
            input: [0; 36],
            input_reads: [[0; 5]; 36],
            input_length_read: read_words(&event.input_length_record),
            output_writes: [[0; 6]; 16],
            pad: [0; 9],
        };
        // xored_state_list[i] is [u64; 25]: lane k = words 2k (low), 2k + 1 (high)
        for (k, lane) in event.xored_state_list[i].iter().enumerate() {
            b.xored_state[2 * k] = *lane as u32;
            b.xored_state[2 * k + 1] = (*lane >> 32) as u32;
        }
        for j in 0..KECCAK_RATE_U32S {
            b.input[j] = event.input[i * KECCAK_RATE_U32S + j];
            b.input_reads[j] = read_words(&event.input_read_records[i * KECCAK_RATE_U32S + j]);
        }
        for j in 0..KECCAK_OUTPUT_U32S {
            b.output_writes[j] = write_words(&event.output_write_records[j]);
        }
        out.push(b);
    }
}

/// Rows of the table: 24 per block, padded to a power of two (trace.rs:86), or the shape's fixed height.
pub fn keccak_sponge_log_height(n_blocks: usize, fixed_log2_rows: Option<usize>) -> usize {
    let rows = n_blocks * KECCAK_ROUNDS;
    match fixed_log2_rows {
        Some(l) => { assert!(rows <= 1 << l, "fixed log2 rows is too small"); l }
        None => rows.next_power_of_two().trailing_zeros() as usize,
    }
}

/// A chip's event vector as it lies in the record: `n_events` records of `words_per_event` 32-bit words.
pub struct EventVector {
    pub words: *const u32,
    pub n_events: usize,
    pub words_per_event: usize,
    pub width: usize,
}

fn vector_of<T>(v: &[T], width: usize) -> Option<EventVector> {
    debug_assert!(core::mem::size_of::<T>() % 4 == 0 && core::mem::align_of::<T>() == 4);
    Some(EventVector { words: v.as_ptr() as *const u32, n_events: v.len(), words_per_event: core::mem::size_of::<T>() / 4, width })
}

// the record sizes libzkb200's row fillers read (csrc/tracegen.cuh alu_event_words)
const _: () = assert!(core::mem::size_of::<zkm_core_executor::events::AluEvent>() == 28);
const _: () = assert!(core::mem::size_of::<zkm_core_executor::events::BranchEvent>() == 28);
const _: () = assert!(core::mem::size_of::<zkm_core_executor::events::JumpEvent>() == 28);
const _: () = assert!(core::mem::size_of::<zkm_core_executor::events::MovCondEvent>() == 28);
const _: () = assert!(core::mem::size_of::<zkm_core_executor::events::CompAluEvent>() == 64);
const _: () = assert!(core::mem::size_of::<zkm_core_executor::events::MemInstrEvent>() == 64);
const _: () = assert!(core::mem::size_of::<zkm_core_executor::events::MiscEvent>() == 60);
const _: () = assert!(core::mem::size_of::<zkm_core_executor::events::SyscallEvent>() == 56);
const _: () = assert!(core::mem::size_of::<zkm_core_executor::events::GlobalLookupEvent>() == 32);
const _: () = assert!(core::mem::size_of::<zkm_core_executor::events::MemoryInitializeFinalizeEvent>() == 16);

/// Chip name (`MachineAir::name`) -> the record field its `generate_trace` walks, with the chip's column count
/// (csrc/tracegen.cuh `alu_width`).
pub fn event_vector(record: &ExecutionRecord, chip: &str) -> Option<EventVector> {
    match chip {
        "AddSub" => vector_of(&record.add_sub_events, 19),
        "Bitwise" => vector_of(&record.bitwise_events, 18),
        "Lt" => vector_of(&record.lt_events, 32),
        "ShiftLeft" => vector_of(&record.shift_left_events, 44),
        "ShiftRight" => vector_of(&record.shift_right_events, 67),
        "CloClz" => vector_of(&record.cloclz_events, 17),
        "Branch" => vector_of(&record.branch_events, 62),
        "Jump" => vector_of(&record.jump_events, 66),
        "MovCond" => vector_of(&record.movcond_events, 32),
        "Mul" => vector_of(&record.mul_events, 58),
        "MemoryInstrs" => vector_of(&record.memory_instr_events, 79),
        "MiscInstrs" => vector_of(&record.misc_events, 72),
        "DivRem" => vector_of(&record.divrem_events, 106),
        "SyscallInstrs" => vector_of(&record.syscall_events, 77),
        // the one table that is not row-local: the library lifts every message to its curve point and scans the points
        "Global" => vector_of(&record.global_lookup_events, 99),
        _ => None,
    }
}

/// Event records the shim has to build (the record does not hold them in the form the row filler reads): `words` holds
/// `n_events` records of `words.len() / n_events` 32-bit words.
pub struct OwnedEvents {
    pub words: Vec<u32>,
    pub n_events: usize,
    pub width: usize,
}

fn syscall_words(e: &zkm_core_executor::events::SyscallEvent) -> [u32; 14] {
    let r = &e.a_record;
    [e.pc, e.next_pc, e.shard, e.clk, r.value, r.shard, r.timestamp, r.prev_value, r.prev_shard, r.prev_timestamp,
     e.a_record_is_real as u32, e.syscall_id, e.arg1, e.arg2]
}

/// * "SyscallCore": the events `SyscallChip::generate_trace` keeps (crates/core/machine/src/syscall/chip.rs:233-240).
/// * "SyscallPrecompile": one `SyscallEvent` per precompile event, its `a_record` carrying what `row_fn` takes from the
///   `PrecompileEvent` (chip.rs:207-222) in the convention of include/syscall.hpp `precompile_event_to_row`.
/// * "MemoryGlobalInit" / "MemoryGlobalFinalize": `zkb200_memory_global_event` records - the events sorted by address
///   (crates/core/machine/src/memory/global.rs:130), each with the address its row is compared with (the previous event's,
///   for the first one the public values' previous address) and its position, so that the sequential loop of
///   global.rs:150-180 becomes one independent record per row.
pub fn owned_events(record: &ExecutionRecord, chip: &str) -> Option<OwnedEvents> {
    use zkm_core_executor::events::PrecompileEvent;
    match chip {
        "SyscallCore" => {
            let kept: Vec<[u32; 14]> = record.syscall_events.iter()
                .filter(|e| { let b = e.a_record.prev_value.to_le_bytes(); b[2] == 1 || b[1] != 0 })
                .map(syscall_words).collect();
            Some(OwnedEvents { n_events: kept.len(), words: kept.concat(), width: 11 })
        }
        "SyscallPrecompile" => {
            let all: Vec<[u32; 14]> = record.precompile_events.all_events().map(|(e, p)| {
                let mut w = syscall_words(e);
                match p {
                    PrecompileEvent::Linux(l) => { w[4] = l.v0; w[7] = 1; }
                    _ => { w[4] = 0; w[7] = 0; }
                }
                w
            }).collect();
            Some(OwnedEvents { n_events: all.len(), words: all.concat(), width: 11 })
        }
        "MemoryGlobalInit" | "MemoryGlobalFinalize" => {
            let init = chip == "MemoryGlobalInit";
            let mut ev = if init { record.global_memory_initialize_events.clone() } else { record.global_memory_finalize_events.clone() };
            let bits = if init { record.public_values.previous_init_addr_bits } else { record.public_values.previous_finalize_addr_bits };
            let mut prev: u32 = bits.iter().enumerate().map(|(j, b)| b << j).sum();
            ev.sort_by_key(|e| e.addr);
            let n = ev.len();
            let mut words = Vec::with_capacity(6 * n);
            for (i, e) in ev.iter().enumerate() {
                let position = (i == 0) as u32 | ((i + 1 == n) as u32) << 1;
                words.extend_from_slice(&[e.addr, e.value, e.shard, e.timestamp, prev, position]);
                prev = e.addr;
            }
            Some(OwnedEvents { n_events: n, words, width: 111 })
        }
        _ => None,
    }
}

/// NUM_LOCAL_MEMORY_ENTRIES_PER_ROW (crates/core/machine/src/memory/local.rs:26)
pub const MEMORY_LOCAL_ENTRIES_PER_ROW: usize = 4;

/// `next_power_of_two(n, fixed_log2_rows)` of crates/core/machine/src/utils/mod.rs:101-125 (at least 16 rows).
pub fn padded_height(n_events: usize, fixed_log2_rows: Option<usize>) -> usize {
    match fixed_log2_rows {
        Some(l) => { assert!(n_events <= 1 << l, "fixed log2 rows is too small"); 1 << l }
        None => n_events.next_power_of_two().max(16),
    }
}

/// `zkb200_cpu_event` (include/zkb200.h): 28 words.
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct CpuEventFlat {
    pub clk: u32,
    pub pc: u32,
    pub next_pc: u32,
    pub next_next_pc: u32,
    pub a: u32,
    pub b: u32,
    pub c: u32,
    pub hi: u32,
    pub flags: u32,
    pub op_word: u32,
    pub op_b: u32,
    pub op_c: u32,
    pub a_record: [u32; 6],
    pub b_record: [u32; 5],
    pub c_record: [u32; 5],
}
const _: () = assert!(core::mem::size_of::<CpuEventFlat>() == 28 * 4);

/// What `CpuChip::generate_trace` reads per row (cpu/trace.rs:45-75): the event, `input.program.fetch(event.pc)` and
/// `input.public_values.execution_shard`.
pub fn flatten_cpu_events(record: &ExecutionRecord) -> Vec<CpuEventFlat> {
    use zkm_core_executor::events::MemoryRecordEnum;
    let shard = record.public_values.execution_shard;
    debug_assert!(shard < 1 << 16, "the Cpu chip range-checks the shard number to 16 bits");
    record.cpu_events.iter().map(|e| {
        let ins = record.program.fetch(e.pc);
        let mut f = CpuEventFlat {
            clk: e.clk, pc: e.pc, next_pc: e.next_pc, next_next_pc: e.next_next_pc, a: e.a, b: e.b, c: e.c,
            hi: e.hi.unwrap_or(0),
            op_word: ins.opcode as u32 | (ins.op_a as u32) << 8 | shard << 16,
            op_b: ins.op_b, op_c: ins.op_c,
            ..Default::default()
        };
        f.flags = e.hi.is_some() as u32 | (ins.imm_b as u32) << 5 | (ins.imm_c as u32) << 6;
        match e.a_record {
            Some(MemoryRecordEnum::Read(r)) => { f.flags |= 1 << 1; f.a_record[..5].copy_from_slice(&read_words(&r)); }
            Some(MemoryRecordEnum::Write(w)) => { f.flags |= 2 << 1; f.a_record = write_words(&w); }
            None => {}
        }
        if let Some(MemoryRecordEnum::Read(r)) = e.b_record { f.flags |= 1 << 3; f.b_record = read_words(&r); }
        if let Some(MemoryRecordEnum::Read(r)) = e.c_record { f.flags |= 1 << 4; f.c_record = read_words(&r); }
        f
    }).collect()
}

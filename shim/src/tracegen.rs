//! Trace generation moved to the device: for the chips libzkb200 has a row filler for,
//! the host hands `zkb200_commit` the record's EVENT VECTORS instead of generated rows (`ZKB200_TRACE_EVENTS`,
//! include/zkb200.h).  The table is written column-major straight into the shard's trace storage on the GPU: none of its
//! bytes cross PCIe and no layout change runs (KeccakSponge: 1.5 KB of record per 24 rows x 3531 columns = 339 KB of rows).
//!
//! * ALU / control-flow chips: `AluEvent`, `BranchEvent`, `JumpEvent`, `MovCondEvent` are `#[repr(C)]` seven-word records
//!   (crates/core/executor/src/events/instr.rs) and cross as they lie in `record.add_events` etc.
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

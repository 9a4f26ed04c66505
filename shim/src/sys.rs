//! Raw bindings of `include/zkb200.h` (the C ABI of libzkb200.so).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct zkb200_trace {
    pub name: *const c_char,
    pub data: *const u32, // row-major height x width, Montgomery: RowMajorMatrix<KoalaBear>::values as is
    pub height: usize,
    pub width: usize,
    pub flags: u32,       // 0, ZKB200_TRACE_COL_MAJOR (1) or ZKB200_TRACE_EVENTS (2)
    pub n_events: usize,  // ZKB200_TRACE_EVENTS: event records behind `data`
}
pub const ZKB200_TRACE_COL_MAJOR: u32 = 1;
pub const ZKB200_TRACE_EVENTS: u32 = 2;
pub const ZKB200_TRACE_DERIVED: u32 = 4; // zkb200_prove_shard only: Byte / Program multiplicities counted on the device (K7)
/// A table resident on the device, column-major Montgomery (`zkb200_table`): a sender of `zkb200_derive_multiplicities`.
#[repr(C)]
pub struct zkb200_table {
    pub chip: *const c_char,
    pub prep: *const u32,       // null for a chip without preprocessed columns
    pub main_trace: *const u32,
    pub height: usize,
}
pub enum zkb200_ctx {}
pub enum zkb200_pk {}
pub enum zkb200_shard {}

extern "C" {
    pub fn zkb200_ctx_create(device: c_int, desc: *const u32, n_words: usize, out: *mut *mut zkb200_ctx) -> c_int;
    pub fn zkb200_ctx_create_multi(devices: *const c_int, n_devices: c_int, desc: *const u32, n_words: usize,
                                   out: *mut *mut zkb200_ctx) -> c_int;
    pub fn zkb200_ctx_num_devices(ctx: *const zkb200_ctx) -> c_int;
    pub fn zkb200_ctx_destroy(ctx: *mut zkb200_ctx);
    pub fn zkb200_last_error(ctx: *mut zkb200_ctx) -> *const c_char;
    pub fn zkb200_setup(ctx: *mut zkb200_ctx, prep: *const zkb200_trace, n: c_int, pc_start: u32,
                        init_global_sum: *const u32, commit_out: *mut u32, out: *mut *mut zkb200_pk) -> c_int;
    pub fn zkb200_pk_free(pk: *mut zkb200_pk);
    pub fn zkb200_pk_initial_challenger(pk: *const zkb200_pk, challenger: *mut u32) -> c_int;
    pub fn zkb200_commit(ctx: *mut zkb200_ctx, traces: *const zkb200_trace, n: c_int, public_values: *const u32,
                         n_public_values: usize, commit_out: *mut u32, out: *mut *mut zkb200_shard) -> c_int;
    pub fn zkb200_shard_free(shard: *mut zkb200_shard);
    pub fn zkb200_open(ctx: *mut zkb200_ctx, pk: *const zkb200_pk, shard: *mut zkb200_shard, challenger: *mut u32,
                       proof_words: *mut *mut u32, n_words: *mut usize) -> c_int;
    pub fn zkb200_free(p: *mut c_void);
    /// MachineAir::generate_trace of the KeccakSponge chip on the device (tracegen.rs: KeccakBlock = zkb200_keccak_block)
    pub fn zkb200_keccak_sponge_trace_width() -> c_int;
    pub fn zkb200_generate_keccak_sponge_trace(ctx: *mut zkb200_ctx, blocks: *const crate::tracegen::KeccakBlock, n_blocks: usize,
                                               log_height: std::os::raw::c_uint, out: *mut u32, col_major: c_int) -> c_int;
    /// K7: the multiplicity columns of a receive-only table (Byte, Program) from the device-resident rows of its senders -
    /// what generate_dependencies -> record.byte_lookups -> ByteChip::generate_trace (bytes/trace.rs:46-67) and
    /// ProgramChip::generate_trace (program/mod.rs:115-158) count on the host.  Every pointer is device memory.
    pub fn zkb200_derive_multiplicities(ctx: *mut zkb200_ctx, receiver: *const c_char, receiver_prep: *const u32,
                                        receiver_height: usize, senders: *const zkb200_table, n_senders: c_int,
                                        out: *mut u32, n_lookups_out: *mut u64) -> c_int;
}

/* zkb200 — C ABI of the B200 STARK shard prover (libzkb200.so).
 *
 * Drop-in boundary for ONE path of ProjectZKM/Ziren: `MachineProver::{setup, commit, open}` for
 * `SC = KoalaBearPoseidon2` (trait: crates/stark/src/prover.rs:30-184; CPU implementation it
 * replaces: crates/stark/src/prover.rs:202-694; selected through ZKMProverComponents,
 * crates/prover/src/components.rs:6-35).  Host Rust keeps MIPS execution, trace generation and
 * shard scheduling and calls through this header; INTEGRATION.md shows the `extern "C"` block and
 * the `impl MachineProver for B200Prover` shim.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success, non-zero on error, and
 *     `zkb200_last_error(ctx)` describes the failure (MachineProver::Error).
 *   - TRACE MATRICES are row-major `height x width` arrays of uint32 in MONTGOMERY form
 *     (R = 2^32, p = 2^31 - 2^24 + 1): the in-memory layout of `RowMajorMatrix<KoalaBear>`, so the
 *     Rust side passes `trace.values.as_ptr()` without conversion.  Pointers may be host
 *     (pageable or pinned) or device memory; the library detects which.
 *   - everything SMALL crosses in CANONICAL form (residues 0..p-1): commitments, public values,
 *     challenger state, pc_start, cumulative sums and the proof ("ZKPF").
 *   - a context is bound to one GPU; calls on one context are serialised internally, so commit/open
 *     may be invoked from several host threads (prove.rs:487-521 does).
 *
 * Machine descriptor "ZKMD" (uint32 words) — chips as data, the export of a SymbolicAirBuilder
 * walk (crates/stark/src/machine.rs:377-389) plus each chip's sends/receives
 * (crates/stark/src/lookup/lookup.rs:10-19):
 *   0x444d4b5a, 1, n_chips, num_pv_elts, log_blowup, num_queries, pow_bits, then per chip:
 *     name_len, name bytes (LE packed, padded to words),
 *     prep_width, main_width, log_quotient_degree, local_only, commit_scope_is_global,
 *     n_sends, n_receives, n_nodes, n_constraints,
 *     lookups (sends then receives): kind, scope (0 local / 1 global), n_values,
 *         multiplicity VPC, n_values x VPC;  VPC = constant, n_terms, n_terms x (is_main, column, weight)
 *     nodes: n_nodes x (op, a, b), topologically ordered, ops:
 *         0 CONST(a)  1 MAIN(col a, row offset b)  2 PREP(col a, row offset b)  3 PUBLIC(a)
 *         4 IS_FIRST_ROW  5 IS_LAST_ROW  6 IS_TRANSITION  7 ADD  8 SUB  9 MUL  10 NEG
 *     constraints: n_constraints x node id, in `assert_zero` order.
 *   The LogUp constraints (crates/stark/src/permutation.rs:205-347) are derived from the lookups.
 *
 * Proof encoding "ZKPF" (uint32 words, canonical) — the fields of ShardProof
 * (crates/stark/src/types.rs:76-83) and of Plonky3's FriProof:
 *   0x46504b5a, 1, main_commit[8], permutation_commit[8], quotient_commit[8],
 *   n_chips, per chip in shard order (height desc, name asc = chip_ordering):
 *       name, log_degree, prep_w, main_w, perm_w (= 4E), n_quotient_chunks,
 *       preprocessed local[prep_w] next[prep_w], main local/next, permutation local/next  (EF = 4 words),
 *       quotient[n_chunks][4] EF, global_cumulative_sum[14], local_cumulative_sum EF
 *   n_public_values, public_values[]
 *   n_commit_phase_commits, commits[8] each, final_poly EF, pow_witness, n_queries, per query:
 *       n_rounds, per round: n_matrices, per matrix: width, opened row[width]; path_len, path[8] each
 *       n_layers, per layer: sibling_value EF, path_len, path[8] each
 *   pow_witness is the SMALLEST valid witness (the reference takes any: rayon find_any).
 *
 * Challenger image (34 words, canonical): sponge_state[16], n_inputs, input_buffer[8],
 *   n_outputs, output_buffer[8] — the fields of DuplexChallenger<Val, Perm, 16, 8>.
 */
#ifndef ZKB200_H
#define ZKB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zkb200_ctx zkb200_ctx;     /* MachineProver instance on one GPU */
typedef struct zkb200_pk zkb200_pk;       /* MachineProver::DeviceProvingKey */
typedef struct zkb200_shard zkb200_shard; /* ShardMainData<SC, DeviceMatrix, DeviceProverData> */

typedef struct {
  const char* name;      /* chip name (MachineAir::name) */
  const uint32_t* data;  /* row-major height x width, Montgomery; host or device pointer (flags = 0) */
  size_t height;         /* power of two */
  size_t width;
  uint32_t flags;        /* 0, or one of ZKB200_TRACE_* (zkb200_commit / zkb200_prove_shard only) */
  size_t n_events;       /* ZKB200_TRACE_EVENTS: number of event records behind `data` */
} zkb200_trace;
/* `data` is a DEVICE pointer to a COLUMN-MAJOR matrix (what zkb200_generate_*_trace(col_major = 1) writes): the
 * row-major -> column-major layout change of the commit is skipped. */
#define ZKB200_TRACE_COL_MAJOR 1u
/* MachineAir::generate_trace moved into the commit: `data` points to the chip's EVENT
 * RECORDS (host or device memory; n_events of them: the 28-byte records of zkb200_generate_alu_trace, or
 * zkb200_keccak_block records for "KeccakSponge"), height = the padded number of rows, width = the chip's width.
 * The library's row filler writes the table column-major straight into the shard's trace storage; no trace bytes
 * cross PCIe and no layout change runs.  Error if the library has no row filler for `name`. */
#define ZKB200_TRACE_EVENTS 2u
/* A table that only RECEIVES lookups (Byte, Program): nothing is handed over (`data` is ignored, height = the height of its
 * preprocessed trace); its multiplicity columns are counted on the device from the shard's other tables (K7,
 * zkb200_derive_multiplicities below) - what the host accumulates in the reference while it generates those tables
 * (crates/core/machine/src/bytes/trace.rs:46-67, program/mod.rs:115-158).  zkb200_prove_shard only: the proving key holds the
 * preprocessed tables the count needs; zkb200_commit refuses the flag.  Excludes the other flags. */
#define ZKB200_TRACE_DERIVED 4u

/* MachineProver::new(machine) — crates/stark/src/prover.rs:43.  `desc` is a ZKMD descriptor. */
int zkb200_ctx_create(int device, const uint32_t* desc, size_t n_words, zkb200_ctx** out);
/* One prover object over several GPUs of the node (n_devices <= 0, or device = -1 above: every visible
 * GPU).  prove_with_context hands ONE MachineProver to all its worker threads
 * (crates/core/machine/src/utils/prove.rs:487-521): zkb200_commit routes each shard to the device with
 * the fewest shards in flight (or to the device that already holds device-resident traces),
 * zkb200_open follows the shard, zkb200_setup replicates the proving key on every device. */
int zkb200_ctx_create_multi(const int* devices, int n_devices, const uint32_t* desc, size_t n_words, zkb200_ctx** out);
int zkb200_ctx_num_devices(const zkb200_ctx* ctx);
/* process-wide experiment knobs used by tools/ (A/B timing inside one process): "ntt_k2" (size of the
 * contiguous NTT level, 0 = built-in), "eval_v2" (0/1, -1 = default). */
int zkb200_set_option(const char* key, long value);
/* The constraint kernels (K3) are generated per chip from the ZKMD descriptor and compiled with NVRTC at run
 * time.  This entry point does only that (no GPU needed): chips compiled, or -1; total cubin bytes out. */
void zkb200_quotient_launch_counts(unsigned long long* generated, unsigned long long* interpreted);
int zkb200_codegen_compile_check(const uint32_t* desc, size_t n_words, size_t* bytes_out);
/* tools/h2d_probe.py: time one pass of a pinned row-major `rows x row_bytes` buffer over PCIe.  mode 0: 2-D DMA
 * in column slices of seg_bytes over n concurrent streams; 1: the pull kernel with n CTAs; 2: contiguous DMA in n parts. */
int zkb200_h2d_probe(zkb200_ctx* ctx, int mode, size_t row_bytes, size_t rows, size_t seg_bytes, int n, float* ms_out);
void zkb200_ctx_destroy(zkb200_ctx* ctx);
/* error text of the last failed call made by the calling host thread (ctx may be NULL) */
const char* zkb200_last_error(zkb200_ctx* ctx);
/* the CUDA stream (cudaStream_t) all work of this context is issued on */
void* zkb200_ctx_stream(zkb200_ctx* ctx);

/* MachineProver::setup + pk_to_device — prover.rs:49-63, machine.rs:352-459: commit the
 * preprocessed traces, keep traces + LDEs + tree on the device.  init_global_sum: 14 canonical
 * words (x[7], y[7]); NULL = SepticDigest::zero(), the curve START point (septic_digest.rs:9-42). */
int zkb200_setup(zkb200_ctx* ctx, const zkb200_trace* prep, int n, uint32_t pc_start,
                 const uint32_t* init_global_sum, uint32_t commit_out[8], zkb200_pk** out);
void zkb200_pk_free(zkb200_pk* pk);
/* MachineProvingKey::observe_into on a fresh challenger — prover.rs:714-721 */
int zkb200_pk_initial_challenger(const zkb200_pk* pk, uint32_t challenger[34]);

/* MachineProver::commit — prover.rs:258-292: sort by (height desc, name), LDE + Merkle. */
int zkb200_commit(zkb200_ctx* ctx, const zkb200_trace* traces, int n, const uint32_t* public_values,
                  size_t n_public_values, uint32_t commit_out[8], zkb200_shard** out);
void zkb200_shard_free(zkb200_shard* shard);
int zkb200_shard_device(const zkb200_shard* shard);   /* CUDA device the shard was committed on */

/* MachineProver::open — prover.rs:298-653.  `challenger` (in/out) is the per-shard clone of the
 * post-observe_into challenger.  The proof buffer is malloc'ed; release with zkb200_free. */
int zkb200_open(zkb200_ctx* ctx, const zkb200_pk* pk, zkb200_shard* shard, uint32_t challenger[34],
                uint32_t** proof_words, size_t* n_words);
/* commit + open of one record (the body of MachineProver::prove's loop, prover.rs:681-688) */
int zkb200_prove_shard(zkb200_ctx* ctx, const zkb200_pk* pk, const zkb200_trace* traces, int n,
                       const uint32_t* public_values, size_t n_public_values, uint32_t challenger[34],
                       uint32_t** proof_words, size_t* n_words);
void zkb200_free(void* p);

/* per-stage device times (ms) of the last zkb200_open when profiling is on; returns the number of
 * stages; names/ms may be NULL to query the count */
void zkb200_set_profile(zkb200_ctx* ctx, int on);
int zkb200_last_stage_times(zkb200_ctx* ctx, const char** names, float* ms, int cap);
/* number of CUDA kernels this library has launched in this process (all contexts) */
unsigned long long zkb200_launch_count(void);

/* ---- kernel-level entry points (tests, micro-benchmarks, ncu captures) -------------------------
 * All matrices here are DEVICE pointers, COLUMN-MAJOR (column c at data + c*height), Montgomery. */
/* K1: out (n<<log_blowup) x width = evaluations on shift*K, rows bit-reversed */
int zkb200_coset_lde(zkb200_ctx* ctx, const uint32_t* in, uint32_t* out, unsigned log_n, size_t width,
                     unsigned log_blowup, uint32_t shift_canonical);
/* K1: plain DFT, natural order in; bitrev_out selects the output row order */
int zkb200_ntt(zkb200_ctx* ctx, const uint32_t* in, uint32_t* out, unsigned log_n, size_t width, int inverse,
               int bitrev_out);
/* K2: MMCS root (canonical) over matrices as given */
int zkb200_mmcs_root(zkb200_ctx* ctx, const uint32_t* const* mats, const unsigned* log_heights, const size_t* widths,
                     int n, uint32_t root_out[8]);
/* K2: n Poseidon2 permutations in place, states row-major n x 16 */
int zkb200_poseidon2_permute_batch(zkb200_ctx* ctx, uint32_t* states, size_t n);
/* K7: the multiplicity columns of a table that only RECEIVES lookups (Byte, Program, range tables), derived from the rows of
 * the tables that send to it.  In the reference these are histograms the host accumulates while it generates every other
 * table: each chip's generate_dependencies -> record.byte_lookups -> ByteChip::generate_trace
 * (crates/core/machine/src/bytes/trace.rs:46-67), the Cpu events' pcs -> ProgramChip::generate_trace
 * (crates/core/machine/src/program/mod.rs:115-158).  Here the machine description already says which tuple every row of every
 * chip sends (the sends of its Air::eval, as in `Chip::sends`, crates/stark/src/chip.rs), so ONE pass over the resident rows -
 * uploaded or generated on the device - gives the same multiplicities without the per-chip byte-lookup bookkeeping.
 * `receiver`: a chip with receives whose values are made of preprocessed columns only and whose multiplicity is a main column;
 * `receiver_prep`: its preprocessed trace; `senders`: the shard's other tables (chips without a matching send are skipped);
 * every pointer is DEVICE memory, column-major, Montgomery.  `out`: receiver_height x main_width of the receiver, column-major
 * Montgomery (columns that are not multiplicities are zero).  Among equal tuples of the receiver the lowest (receive, row)
 * takes the multiplicity.  Fails (MachineProver::Error) when a sent tuple is in no row of the receiver. */
typedef struct {
  const char* chip;
  const uint32_t* prep;         /* NULL for a chip without preprocessed columns */
  const uint32_t* main_trace;
  size_t height;
} zkb200_table;
int zkb200_derive_multiplicities(zkb200_ctx* ctx, const char* receiver, const uint32_t* receiver_prep, size_t receiver_height,
                                 const zkb200_table* senders, int n_senders, uint32_t* out, uint64_t* n_lookups_out);
/* K5: LogUp permutation trace of one chip; out n x 4E column-major, local_sum canonical[4] */
int zkb200_permutation_trace(zkb200_ctx* ctx, const char* chip, const uint32_t* prep, const uint32_t* main_trace,
                             size_t height, const uint32_t alpha[4], const uint32_t beta[4], uint32_t* out,
                             uint32_t local_sum_out[4]);
/* K3: quotient chunks of one chip from committed LDEs; challenges/sums canonical; pub on host.
 * out: 2^lqd chunk matrices, each n x 4 column-major */
int zkb200_quotient(zkb200_ctx* ctx, const char* chip, unsigned log_n, const uint32_t* prep_lde,
                    const uint32_t* main_lde, const uint32_t* perm_lde, const uint32_t perm_alpha[4],
                    const uint32_t perm_beta[4], const uint32_t local_sum[4], const uint32_t global_sum[14],
                    const uint32_t alpha[4], const uint32_t* public_values, size_t n_public_values, uint32_t* out);
/* K4c: one FRI fold of m EF values (component-major [4][m]); ro_next may be NULL */
int zkb200_fri_fold(zkb200_ctx* ctx, const uint32_t* in, size_t m, const uint32_t beta[4], const uint32_t* ro_next,
                    uint32_t* out);
/* K4d: smallest proof-of-work witness for a challenger image */
int zkb200_grind(zkb200_ctx* ctx, const uint32_t challenger[34], unsigned bits, uint32_t* witness_out);
/* ---- trace generation ---------------------------------------------
 * MachineAir::generate_trace of the core ALU chips AddSub, Bitwise, Lt, ShiftLeft, ShiftRight, CloClz
 * (crates/core/machine/src/alu/{add_sub,bitwise,lt,sll,sr,clo_clz}/mod.rs) and of the control-flow
 * chips Branch and Jump (crates/core/machine/src/control_flow/{branch,jump}/trace.rs) and of MovCond
 * (crates/core/machine/src/misc/mov_cond/mod.rs); the reference's
 * own C++ twins are the chip headers under crates/core/machine/include/ behind cpp/extern.cpp:15-90.  One event per row
 * in event order, then the chip's padding rows up to 2^log_height (next_power_of_two /
 * fixed_log2_rows, crates/core/machine/src/utils/mod.rs:101-125, is the caller's choice).
 * `events` is the record's event vector as it lies in memory, host or device: 28-byte #[repr(C)]
 * records, `AluEvent` {pc, next_pc, opcode, hi, a, b, c} for the ALU chips and `BranchEvent` /
 * `JumpEvent` {pc, next_pc, next_next_pc, opcode, a, b, c} for Branch / Jump, `MovCondEvent`
 * {pc, next_pc, opcode, a, b, c, prev_a} for MovCond
 * (crates/core/executor/src/events/instr.rs:11-26, :160-217, :287-302).  "Mul" (crates/core/machine/src/alu/mul/mod.rs, C++
 * twin include/mul.hpp) takes 64-byte `CompAluEvent` records {shard, clk, pc, next_pc, opcode, hi, a, b, c, hi_record
 * {value, shard, timestamp, prev_value, prev_shard, prev_timestamp}, hi_record_is_real} (instr.rs:47-73).  "MemoryInstrs"
 * (crates/core/machine/src/memory/instructions/trace.rs:103-263, C++ twin include/memory_instrs.hpp; 79 columns) takes 64-byte
 * `MemInstrEvent` records {shard, clk, pc, next_pc, opcode, a, b, c, mem_access {tag: 0 Read / 1 Write, record: six words},
 * prev_a_val} (instr.rs:114-136, events/memory.rs:46-97).  "MemoryLocal" (crates/core/machine/src/memory/local.rs:146-190, C++
 * twin include/memory_local.hpp; 56 columns) takes 28-byte `MemoryLocalEvent` records {addr, initial_mem_access {shard,
 * timestamp, value}, final_mem_access {shard, timestamp, value}} (events/memory.rs:228-237) and packs FOUR events into a row:
 * n_events may be up to 4 * 2^log_height.  "MiscInstrs" (crates/core/machine/src/misc/others/trace.rs:91-275, C++ twin
 * include/misc_instrs.hpp; 72 columns, 44 of them a union per opcode family) takes 60-byte `MiscEvent` records {shard, clk, pc,
 * next_pc, opcode, a, b, c, prev_a, hi_record[6]} (instr.rs:241-261).  `out` is DEVICE memory of
 * 2^log_height x width words, Montgomery, row-major (col_major = 0: the RowMajorMatrix layout
 * zkb200_commit takes) or column-major (col_major = 1: the layout of the kernel-level entry points). */
typedef struct {
  uint32_t pc, next_pc;
  uint8_t opcode;        /* Opcode as #[repr(u8)], crates/core/executor/src/opcode.rs:25-89 */
  uint32_t hi, a, b, c;
} zkb200_alu_event;
typedef struct {
  uint32_t pc, next_pc, next_next_pc;
  uint8_t opcode;
  uint32_t a, b, c;
} zkb200_flow_event;     /* BranchEvent / JumpEvent */
/* "Cpu" (crates/core/machine/src/cpu/trace.rs:45-246, C++ twin include/cpu.hpp behind cpu_event_to_row_koalabear(CpuEventFfi,
 * shard, InstructionFfi), cpp/extern.cpp:6-14; 67 columns, cpu/columns/mod.rs:18-84).  CpuEvent
 * (crates/core/executor/src/events/cpu.rs:15-44) holds Options and the instruction is fetched from the program, so the host
 * writes one flat 112-byte record per event holding what the row needs of the event, of `program.fetch(event.pc)` and the
 * shard number:
 *   flags     bit 0: hi is Some; bits 1-2: a_record 0 None / 1 Read / 2 Write; bit 3: b_record is Some(Read); bit 4: c_record
 *             is Some(Read); bit 5: instruction.imm_b; bit 6: instruction.imm_c
 *   op_word   instruction.opcode | instruction.op_a << 8 | public_values.execution_shard << 16
 *   a_record  the MemoryRecordEnum payload as it lies: Read {value, shard, timestamp, prev_shard, prev_timestamp, -},
 *             Write {value, shard, timestamp, prev_value, prev_shard, prev_timestamp} (events/memory.rs:46-82)
 *   b_record, c_record   MemoryReadRecord {value, shard, timestamp, prev_shard, prev_timestamp}
 * Rows past the last event are the chip's padding rows (imm_b = imm_c = is_rw_a = 1, trace.rs:60-63). */
typedef struct {
  uint32_t clk, pc, next_pc, next_next_pc, a, b, c, hi;
  uint32_t flags, op_word, op_b, op_c;
  uint32_t a_record[6], b_record[5], c_record[5];
} zkb200_cpu_event;
/* "DivRem" (crates/core/machine/src/alu/divrem/mod.rs:229-364, C++ twin include/div_rem.hpp; 106 columns) takes 64-byte
 * `CompAluEvent` records as they lie in record.divrem_events, like "Mul".  "SyscallInstrs"
 * (crates/core/machine/src/syscall/instructions/trace.rs:89-177, C++ twin include/syscall_instrs.hpp; 77 columns) takes the
 * 56-byte `SyscallEvent` records {pc, next_pc, shard, clk, a_record[6], a_record_is_real, syscall_id, arg1, arg2}
 * (crates/core/executor/src/events/syscall.rs:8-29) of record.syscall_events.  "SyscallCore" and "SyscallPrecompile"
 * (crates/core/machine/src/syscall/chip.rs:184-268, C++ twin include/syscall.hpp; 11 columns) take `SyscallEvent` records too:
 * SyscallCore the events generate_trace keeps (a_record.prev_value byte 2 == 1 or byte 1 != 0, chip.rs:233-240 - the caller
 * filters, the rows are dense), SyscallPrecompile one event per precompile event whose a_record carries what row_fn reads from
 * the PrecompileEvent, in the convention of syscall.hpp precompile_event_to_row: prev_value = 1 and value = v0 for
 * PrecompileEvent::Linux, prev_value = 0 otherwise.
 * "MemoryGlobalInit" / "MemoryGlobalFinalize" (crates/core/machine/src/memory/global.rs:115-192, C++ twin
 * include/memory_global.hpp for the per-event columns; 111 columns): the `MemoryInitializeFinalizeEvent`s
 * (crates/core/executor/src/events/memory.rs:138-149) SORTED BY ADDRESS as generate_trace sorts them, each followed by the
 * address its row is compared with and by its position - the two things the reference's sequential loop (global.rs:150-180)
 * takes from the neighbouring row, so that every row depends on its own record only. */
/* "Global" (crates/core/machine/src/global/mod.rs:115-194; 99 columns) takes the 32-byte `GlobalLookupEvent` records {message[7],
 * is_receive: bool, kind: u8} (crates/core/executor/src/events/global.rs:6-15) as they lie in record.global_lookup_events.  It is
 * the one table that is not row-local and whose rows are compute-heavy: every message is lifted to a point of the septic curve
 * (SepticCurve::lift_x, crates/stark/src/septic_curve.rs:130-154: a square test and a square root in F_p^7 per trial), the points
 * are summed along the table (a scan under the curve addition) and each row holds two neighbouring sums. */
typedef struct {
  uint32_t message[7];
  uint32_t is_receive_kind;                 /* byte 0: is_receive, byte 1: kind (LookupKind), bytes 2-3 padding */
} zkb200_global_lookup_event;
typedef struct {
  uint32_t addr, value, shard, timestamp;   /* MemoryInitializeFinalizeEvent */
  uint32_t prev_addr;                       /* the previous event's addr; first event: the public values' previous_init_addr /
                                               previous_finalize_addr (their *_addr_bits recombined), 0 in the first shard */
  uint32_t position;                        /* bit 0: first event of the table, bit 1: last event */
} zkb200_memory_global_event;
/* NUM_*_COLS of the chip, -1 if this library has no row filler for it */
int zkb200_alu_trace_width(const char* chip);
/* events: zkb200_alu_event[]; zkb200_flow_event[] for "Branch" / "Jump"; seven-word MovCondEvent records for "MovCond";
 * CompAluEvent / MemInstrEvent / MemoryLocalEvent / MiscEvent records as they lie for "Mul" / "MemoryInstrs" / "MemoryLocal" /
 * "MiscInstrs"; CompAluEvent records for "DivRem"; SyscallEvent records for "SyscallCore" / "SyscallPrecompile" /
 * "SyscallInstrs"; zkb200_memory_global_event[] for "MemoryGlobalInit" / "MemoryGlobalFinalize"; zkb200_global_lookup_event[] for "Global";
 * zkb200_cpu_event[] for "Cpu" */
int zkb200_generate_alu_trace(zkb200_ctx* ctx, const char* chip, const void* events, size_t n_events,
                              unsigned log_height, uint32_t* out, int col_major);
/* MachineAir::generate_trace of the KeccakSponge precompile chip
 * (crates/core/machine/src/syscall/precompiles/keccak_sponge/trace.rs:58-196; the 2633 permutation columns are
 * p3_keccak_air::generate_trace_rows of the pinned Plonky3), 3531 columns (keccak_sponge/columns.rs:14-37).
 * The Rust KeccakSpongeEvent (crates/core/executor/src/events/precompiles/keccak_sponge.rs:15-40) holds Vecs, so
 * the host flattens a record's events into one fixed-size record per absorbed BLOCK (= 24 rows, one per Keccak-f
 * round), event after event, block after block:
 *   `xored_state`      event.xored_state_list[block] as 50 words (u64 lane i = words 2i, 2i+1)
 *   `input`            event.input[36*block .. 36*block+36]
 *   `input_reads`      event.input_read_records[36*block ..], MemoryReadRecord {value, shard, timestamp, prev_shard,
 *                      prev_timestamp} (crates/core/executor/src/events/memory.rs:46-60)
 *   `input_length_read` event.input_length_record (used by block 0), `output_writes` event.output_write_records,
 *                      MemoryWriteRecord {value, shard, timestamp, prev_value, prev_shard, prev_timestamp} (used by
 *                      the last block)
 * Rows past 24 * n_blocks are the chip's dummy rows up to 2^log_height.  `blocks` may be host or device memory;
 * `out` is DEVICE memory, Montgomery, row-major (col_major = 0) or column-major (col_major = 1). */
typedef struct {
  uint32_t shard, clk, input_addr, output_addr;
  uint32_t input_len;      /* event.input.len(), in words */
  uint32_t block;          /* index of this block in its event */
  uint32_t num_blocks;     /* event.num_blocks() */
  uint32_t reserved;
  uint32_t xored_state[50];
  uint32_t input[36];
  uint32_t input_reads[36][5];
  uint32_t input_length_read[5];
  uint32_t output_writes[16][6];
  uint32_t pad[9];
} zkb200_keccak_block;     /* 384 words */
int zkb200_keccak_sponge_trace_width(void);    /* NUM_KECCAK_SPONGE_COLS = 3531 */
int zkb200_generate_keccak_sponge_trace(zkb200_ctx* ctx, const zkb200_keccak_block* blocks, size_t n_blocks,
                                        unsigned log_height, uint32_t* out, int col_major);
/* layout helpers on the context stream: row-major <-> column-major, canonical <-> Montgomery */
int zkb200_transpose(zkb200_ctx* ctx, const uint32_t* in, uint32_t* out, size_t height, size_t width, int to_colmajor);
int zkb200_convert(zkb200_ctx* ctx, uint32_t* data, size_t n, int to_montgomery);
int zkb200_sync(zkb200_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* ZKB200_H */

//! `impl MachineProver<KoalaBearPoseidon2, A> for B200Prover<A>` — the Rust half of the drop-in boundary
//! (trait: crates/stark/src/prover.rs:30-184).  Everything numeric happens behind `include/zkb200.h`; this
//! crate exports the machine as data once (`export::export_zkmd`), marshals traces and the challenger across
//! the C ABI and repacks the flat proof into `ShardProof` (`proof::decode_zkpf`).
//!
//! Select it like any other prover, through `ZKMProverComponents` (crates/prover/src/components.rs:6-35):
//! ```ignore
//! pub struct B200ProverComponents;
//! impl ZKMProverComponents for B200ProverComponents {
//!     type CoreProver = zkm_b200::B200Prover<MipsAir<KoalaBear>>;
//!     type CompressProver = zkm_b200::B200Prover<RecursionAir<KoalaBear, COMPRESS_DEGREE>>;
//!     type ShrinkProver = CpuProver<InnerSC, ShrinkAir<KoalaBear>>;
//!     type WrapProver = CpuProver<OuterSC, WrapAir<KoalaBear>>;       // BN254 Poseidon2: not this library
//! }
//! ```
//! NOT compiled in this repository (no Rust toolchain in the build image); see shim/Cargo.toml.
pub mod export;
pub mod proof;
pub mod sys;
pub mod tracegen;

use std::ffi::{CStr, CString};
use std::ptr::null_mut;

use hashbrown::HashMap;
use p3_air::Air;
use p3_challenger::DuplexChallenger;
use p3_field::{FieldAlgebra, PrimeField32};
use p3_koala_bear::KoalaBear;
use p3_matrix::{dense::RowMajorMatrix, Matrix};
use p3_uni_stark::SymbolicAirBuilder;
use zkm_stark::{
    air::MachineAir, koala_bear_poseidon2::KoalaBearPoseidon2, septic_digest::SepticDigest, Com, CpuProver,
    DebugConstraintBuilder, MachineProof, MachineProver, MachineProvingKey, MachineRecord, ShardMainData, ShardProof,
    StarkGenericConfig, StarkMachine, StarkProvingKey, StarkVerifyingKey, Val,
};

type SC = KoalaBearPoseidon2;
type F = KoalaBear;

#[derive(Debug, thiserror::Error)]
#[error("zkb200: {0}")]
pub struct B200Error(pub String);

fn check(ctx: *mut sys::zkb200_ctx, rc: i32) -> Result<(), B200Error> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(sys::zkb200_last_error(ctx)) }.to_string_lossy().into_owned();
    Err(B200Error(msg))
}

/// The traces stay on the device: `ShardMainData.traces` only needs a `Matrix` to name heights and widths.
pub struct B200Matrix { height: usize, width: usize }
impl Matrix<F> for B200Matrix {
    fn width(&self) -> usize { self.width }
    fn height(&self) -> usize { self.height }
    type Row<'a> = core::iter::Empty<F>;
    fn row(&self, _r: usize) -> Self::Row<'_> { unimplemented!("B200Matrix lives on the device") }
}
/// Owns the device-side ShardMainData (traces, LDEs, Merkle tree) between `commit` and `open`.
pub struct B200ProverData(*mut sys::zkb200_shard);
unsafe impl Send for B200ProverData {}
impl Drop for B200ProverData {
    fn drop(&mut self) { unsafe { sys::zkb200_shard_free(self.0) } }
}

/// Host key (what `ZKMProvingKey` serialises, crates/prover/src/lib.rs:293-301) plus its device image.
pub struct B200ProvingKey { host: StarkProvingKey<SC>, dev: *mut sys::zkb200_pk }
unsafe impl Send for B200ProvingKey {}
unsafe impl Sync for B200ProvingKey {}
impl Drop for B200ProvingKey {
    fn drop(&mut self) { unsafe { sys::zkb200_pk_free(self.dev) } }
}
impl MachineProvingKey<SC> for B200ProvingKey {
    fn preprocessed_commit(&self) -> Com<SC> { self.host.commit.clone() }
    fn pc_start(&self) -> Val<SC> { self.host.pc_start }
    fn initial_global_cumulative_sum(&self) -> SepticDigest<Val<SC>> { self.host.initial_global_cumulative_sum }
    fn observe_into(&self, challenger: &mut <SC as StarkGenericConfig>::Challenger) { self.host.observe_into(challenger) }
}

pub struct B200Prover<A> {
    cpu: CpuProver<SC, A>, // setup / pk_from_vk / debug paths and the owner of the StarkMachine
    ctx: *mut sys::zkb200_ctx,
}
unsafe impl<A: Send> Send for B200Prover<A> {}
unsafe impl<A: Sync> Sync for B200Prover<A> {}
impl<A> Drop for B200Prover<A> {
    fn drop(&mut self) { unsafe { sys::zkb200_ctx_destroy(self.ctx) } }
}

fn marshal(names: &[CString], mats: &[&RowMajorMatrix<F>]) -> Vec<sys::zkb200_trace> {
    names.iter().zip(mats).map(|(n, m)| sys::zkb200_trace {
        name: n.as_ptr(),
        // KoalaBear is #[repr(transparent)] over its Montgomery u32: the values cross as they are
        // (the in-tree C++ does the same cast, crates/core/machine/cpp/extern.cpp:12)
        data: m.values.as_ptr() as *const u32,
        height: m.height(),
        width: m.width(),
        flags: 0,
        n_events: 0,
    }).collect()
}

/// 34-word canonical image of DuplexChallenger<KoalaBear, Perm, 16, 8> (include/zkb200.h)
fn challenger_to_words(c: &<SC as StarkGenericConfig>::Challenger) -> [u32; 34] {
    let mut w = [0u32; 34];
    for (i, x) in c.sponge_state.iter().enumerate() { w[i] = x.as_canonical_u32(); }
    w[16] = c.input_buffer.len() as u32;
    for (i, x) in c.input_buffer.iter().enumerate() { w[17 + i] = x.as_canonical_u32(); }
    w[25] = c.output_buffer.len() as u32;
    for (i, x) in c.output_buffer.iter().enumerate() { w[26 + i] = x.as_canonical_u32(); }
    w
}
fn challenger_from_words(c: &mut <SC as StarkGenericConfig>::Challenger, w: &[u32; 34]) {
    for i in 0..16 { c.sponge_state[i] = F::from_canonical_u32(w[i]); }
    c.input_buffer = (0..w[16] as usize).map(|i| F::from_canonical_u32(w[17 + i])).collect();
    c.output_buffer = (0..w[25] as usize).map(|i| F::from_canonical_u32(w[26 + i])).collect();
}

/// What the device-side trace generation needs from a record; `ExecutionRecord` (the core machine) provides it, the
/// recursion records do not (their tables are uploaded as rows).
pub trait DeviceTraceEvents {
    fn keccak_sponge_blocks(&self) -> Option<Vec<tracegen::KeccakBlock>> { None }
    /// The chip's `#[repr(C)]` event vector as it lies in the record (tracegen.rs `EventVector`), for the chips whose
    /// rows are one event each.
    fn event_vector(&self, _chip: &str) -> Option<tracegen::EventVector> { None }
    /// MemoryLocal: the shard's local memory events gathered into one vector (the record keeps them in several places,
    /// `ExecutionRecord::get_local_mem_events`), four to a row.
    fn memory_local_events(&self) -> Option<Vec<zkm_core_executor::events::MemoryLocalEvent>> { None }
    /// Cpu: one `zkb200_cpu_event` per `CpuEvent` with its fetched instruction and the shard number (tracegen.rs).
    fn cpu_events(&self) -> Option<Vec<tracegen::CpuEventFlat>> { None }
    /// SyscallCore, SyscallPrecompile, MemoryGlobalInit, MemoryGlobalFinalize: records built from the record's events
    /// (tracegen.rs `owned_events`).
    fn owned_events(&self, _chip: &str) -> Option<tracegen::OwnedEvents> { None }
    fn fixed_log2_rows_of(&self, _chip: &str) -> Option<usize> { None }
}
impl DeviceTraceEvents for zkm_core_executor::ExecutionRecord {
    fn keccak_sponge_blocks(&self) -> Option<Vec<tracegen::KeccakBlock>> {
        let b = tracegen::flatten_keccak_sponge_events(self);
        if b.is_empty() { None } else { Some(b) }
    }
    fn event_vector(&self, chip: &str) -> Option<tracegen::EventVector> { tracegen::event_vector(self, chip) }
    fn memory_local_events(&self) -> Option<Vec<zkm_core_executor::events::MemoryLocalEvent>> {
        Some(self.get_local_mem_events().copied().collect())
    }
    fn cpu_events(&self) -> Option<Vec<tracegen::CpuEventFlat>> { Some(tracegen::flatten_cpu_events(self)) }
    fn owned_events(&self, chip: &str) -> Option<tracegen::OwnedEvents> { tracegen::owned_events(self, chip) }
    fn fixed_log2_rows_of(&self, chip: &str) -> Option<usize> {
        self.shape.as_ref().and_then(|s| s.inner.get(chip).copied())
    }
}
impl<A: MachineAir<F>> B200Prover<A>
where
    A::Record: DeviceTraceEvents,
{
    fn keccak_blocks_of(&self, record: &A::Record) -> Option<Vec<tracegen::KeccakBlock>> { record.keccak_sponge_blocks() }
    fn fixed_log2_rows(&self, record: &A::Record, chip: &str) -> Option<usize> { record.fixed_log2_rows_of(chip) }
}

impl<A> MachineProver<SC, A> for B200Prover<A>
where
    A: MachineAir<F> + Air<SymbolicAirBuilder<F>> + Send + Sync + 'static,
    A::Record: DeviceTraceEvents,           // implement it (all defaults) for a record type without device row fillers
    CpuProver<SC, A>: MachineProver<SC, A, DeviceProvingKey = StarkProvingKey<SC>>,
{
    type DeviceMatrix = B200Matrix;
    type DeviceProverData = B200ProverData;
    type DeviceProvingKey = B200ProvingKey;
    type Error = B200Error;

    fn new(machine: StarkMachine<SC, A>) -> Self {
        let fri = machine.config().fri_config();
        let desc = export::export_zkmd(&machine, fri.log_blowup as u32, fri.num_queries as u32, fri.proof_of_work_bits as u32);
        let mut ctx = null_mut();
        // device -1: one prover object over every visible GPU; commit() routes shards (prove.rs:487-521 shares
        // ONE prover between its worker threads)
        let rc = unsafe { sys::zkb200_ctx_create(-1, desc.as_ptr(), desc.len(), &mut ctx) };
        check(null_mut(), rc).expect("zkb200_ctx_create");
        Self { cpu: CpuProver::new(machine), ctx }
    }

    fn machine(&self) -> &StarkMachine<SC, A> { self.cpu.machine() }

    fn setup(&self, program: &A::Program) -> (Self::DeviceProvingKey, StarkVerifyingKey<SC>) {
        let (pk, vk) = self.cpu.setup(program);
        (self.pk_to_device(&pk), vk)
    }
    fn pk_from_vk(&self, program: &A::Program, vk: &StarkVerifyingKey<SC>) -> Self::DeviceProvingKey {
        self.pk_to_device(&self.cpu.pk_from_vk(program, vk))
    }

    /// Uploads `pk.traces` and recomputes LDE + tree on the GPU; the recomputed root must equal `pk.commit`
    /// (a free end-to-end parity check of K1 + K2 against Plonky3).
    fn pk_to_device(&self, pk: &StarkProvingKey<SC>) -> Self::DeviceProvingKey {
        let mut by_index: Vec<(&String, &usize)> = pk.chip_ordering.iter().collect();
        by_index.sort_by_key(|(_, i)| **i);
        let names: Vec<CString> = by_index.iter().map(|(n, _)| CString::new(n.as_str()).unwrap()).collect();
        let mats: Vec<&RowMajorMatrix<F>> = pk.traces.iter().collect();
        let t = marshal(&names, &mats);
        let gsum: Vec<u32> = pk.initial_global_cumulative_sum.0.x.0.iter().chain(pk.initial_global_cumulative_sum.0.y.0.iter())
            .map(|v| v.as_canonical_u32()).collect();
        let (mut commit, mut dev) = ([0u32; 8], null_mut());
        let rc = unsafe { sys::zkb200_setup(self.ctx, t.as_ptr(), t.len() as i32, pk.pc_start.as_canonical_u32(), gsum.as_ptr(), commit.as_mut_ptr(), &mut dev) };
        check(self.ctx, rc).expect("zkb200_setup");
        let want: [F; 8] = pk.commit.clone().into();
        assert!(want.iter().zip(commit).all(|(a, b)| a.as_canonical_u32() == b), "GPU preprocessed commitment differs from the host key's");
        B200ProvingKey { host: pk.clone(), dev }
    }
    fn pk_to_host(&self, pk: &Self::DeviceProvingKey) -> StarkProvingKey<SC> { pk.host.clone() }

    fn commit(&self, record: &A::Record, traces: Vec<(String, RowMajorMatrix<Val<SC>>)>) -> ShardMainData<SC, B200Matrix, B200ProverData> {
        let _span = tracing::debug_span!("commit to main traces (b200)").entered();
        let names: Vec<CString> = traces.iter().map(|(n, _)| CString::new(n.as_str()).unwrap()).collect();
        let mats: Vec<&RowMajorMatrix<F>> = traces.iter().map(|(_, m)| m).collect();
        let mut t = marshal(&names, &mats);
        // Trace generation on the device (tracegen.rs): a core-machine caller whose `generate_traces` left the
        // KeccakSponge table out as an EMPTY matrix of the right width hands its event records over here, and
        // libzkb200's row filler writes the table inside the commit (ZKB200_TRACE_EVENTS).
        let keccak_blocks = self.keccak_blocks_of(record);
        if let Some(blocks) = keccak_blocks.as_ref() {
            if let Some(i) = traces.iter().position(|(n, m)| n == "KeccakSponge" && m.values.is_empty()) {
                let log_h = tracegen::keccak_sponge_log_height(blocks.len(), self.fixed_log2_rows(record, "KeccakSponge"));
                t[i].data = blocks.as_ptr() as *const u32;
                t[i].height = 1 << log_h;
                t[i].width = tracegen::NUM_KECCAK_SPONGE_COLS;
                t[i].flags = sys::ZKB200_TRACE_EVENTS;
                t[i].n_events = blocks.len();
            }
        }
        // The one-event-per-row chips: the record's event vector crosses as it lies (28 or 64 bytes per row instead of
        // 4 x width), same convention - the caller left the table out as an empty matrix of the chip's width.
        for (i, (name, m)) in traces.iter().enumerate() {
            if !m.values.is_empty() || name == "KeccakSponge" { continue; }
            if let Some(ev) = record.event_vector(name) {
                assert_eq!(m.width(), ev.width, "row filler of {name} writes another width");
                t[i].data = ev.words;
                t[i].height = tracegen::padded_height(ev.n_events, self.fixed_log2_rows(record, name));
                t[i].flags = sys::ZKB200_TRACE_EVENTS;
                t[i].n_events = ev.n_events;
            }
        }
        // SyscallCore / SyscallPrecompile / MemoryGlobalInit / MemoryGlobalFinalize: records built here, kept alive until the
        // commit returns
        let owned: Vec<(usize, tracegen::OwnedEvents)> = traces.iter().enumerate()
            .filter(|(_, (_, m))| m.values.is_empty())
            .filter_map(|(i, (n, _))| record.owned_events(n).map(|ev| (i, ev))).collect();
        for (i, ev) in owned.iter() {
            assert_eq!(traces[*i].1.width(), ev.width, "row filler of {} writes another width", traces[*i].0);
            t[*i].data = ev.words.as_ptr();
            t[*i].height = tracegen::padded_height(ev.n_events, self.fixed_log2_rows(record, &traces[*i].0));
            t[*i].flags = sys::ZKB200_TRACE_EVENTS;
            t[*i].n_events = ev.n_events;
        }
        // MemoryLocal: four seven-word events per row, gathered from the record's local-memory iterators
        let local_events = traces.iter().position(|(n, m)| n == "MemoryLocal" && m.values.is_empty())
            .and_then(|i| record.memory_local_events().map(|v| (i, v)));
        if let Some((i, ev)) = local_events.as_ref() {
            const _: () = assert!(core::mem::size_of::<zkm_core_executor::events::MemoryLocalEvent>() == 28);
            t[*i].data = ev.as_ptr() as *const u32;
            t[*i].height = tracegen::padded_height(ev.len().div_ceil(tracegen::MEMORY_LOCAL_ENTRIES_PER_ROW), self.fixed_log2_rows(record, "MemoryLocal"));
            t[*i].flags = sys::ZKB200_TRACE_EVENTS;
            t[*i].n_events = ev.len();
        }
        // Cpu: 112-byte flat records (event + fetched instruction + shard) instead of 268-byte rows
        let cpu_events = traces.iter().position(|(n, m)| n == "Cpu" && m.values.is_empty())
            .and_then(|i| record.cpu_events().map(|v| (i, v)));
        if let Some((i, ev)) = cpu_events.as_ref() {
            t[*i].data = ev.as_ptr() as *const u32;
            // CpuChip::num_rows (cpu/trace.rs:32-43): the shape's height, else at least 16 rows
            t[*i].height = tracegen::padded_height(ev.len(), self.fixed_log2_rows(record, "Cpu"));
            t[*i].flags = sys::ZKB200_TRACE_EVENTS;
            t[*i].n_events = ev.len();
        }
        let public_values = record.public_values::<F>();
        let pv: Vec<u32> = public_values.iter().map(|v| v.as_canonical_u32()).collect();
        let (mut commit, mut shard) = ([0u32; 8], null_mut());
        let rc = unsafe { sys::zkb200_commit(self.ctx, t.as_ptr(), t.len() as i32, pv.as_ptr(), pv.len(), commit.as_mut_ptr(), &mut shard) };
        check(self.ctx, rc).expect("zkb200_commit");          // the trait's commit is infallible (prover.rs:110-114)
        // chip ordering as CpuProver::commit sorts it: (Reverse(height), name), prover.rs:264
        // heights and widths as handed to the library (a table passed as event records has its padded height there)
        let mut order: Vec<(usize, usize, &str)> = traces.iter().zip(&t).map(|((n, _), d)| (d.height, d.width, n.as_str())).collect();
        order.sort_by(|a, b| b.0.cmp(&a.0).then(a.2.cmp(b.2)));
        let chip_ordering: HashMap<String, usize> = order.iter().enumerate().map(|(i, (_, _, n))| (n.to_string(), i)).collect();
        let shapes = order.iter().map(|(h, w, _)| B200Matrix { height: *h, width: *w }).collect();
        let main_commit: [F; 8] = core::array::from_fn(|i| F::from_canonical_u32(commit[i]));
        ShardMainData::new(shapes, main_commit.into(), B200ProverData(shard), chip_ordering, public_values)
    }

    fn open(&self, pk: &B200ProvingKey, data: ShardMainData<SC, B200Matrix, B200ProverData>,
            challenger: &mut <SC as StarkGenericConfig>::Challenger) -> Result<ShardProof<SC>, B200Error> {
        let _span = tracing::debug_span!("open multi batches (b200)").entered();
        let mut st = challenger_to_words(challenger);
        let (mut words, mut n) = (null_mut(), 0usize);
        let rc = unsafe { sys::zkb200_open(self.ctx, pk.dev, data.main_data.0, st.as_mut_ptr(), &mut words, &mut n) };
        check(self.ctx, rc)?;
        let proof = proof::decode_zkpf(unsafe { std::slice::from_raw_parts(words, n) });
        unsafe { sys::zkb200_free(words as *mut _) };
        challenger_from_words(challenger, &st);
        Ok(proof)
    }

    /// `CpuProver::prove` (prover.rs:656-693) with this prover's commit / open.
    fn prove(&self, pk: &B200ProvingKey, mut records: Vec<A::Record>, challenger: &mut <SC as StarkGenericConfig>::Challenger,
             opts: <A::Record as MachineRecord>::Config) -> Result<MachineProof<SC>, B200Error>
    where
        A: for<'a> Air<DebugConstraintBuilder<'a, Val<SC>, <SC as StarkGenericConfig>::Challenge>>,
    {
        self.machine().generate_dependencies(&mut records, &opts, None);
        pk.observe_into(challenger);
        let shard_proofs = records.iter().map(|record| {
            let traces = self.generate_traces(record).map_err(|e| B200Error(format!("trace generation: {e:?}")))?;
            let data = self.commit(record, traces);
            self.open(pk, data, &mut challenger.clone())
        }).collect::<Result<Vec<_>, _>>()?;
        Ok(MachineProof { shard_proofs })
    }
}

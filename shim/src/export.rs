//! Chips as data: the "ZKMD" machine descriptor of `include/zkb200.h`.
//!
//! The reference monomorphises `chip.eval(&mut folder)` (crates/stark/src/quotient.rs:157); across the C ABI
//! the constraints travel as the DAG a `SymbolicAirBuilder` walk produces - the same walk
//! `StarkMachine::setup` already does to count constraints (crates/stark/src/machine.rs:377-389) - and the
//! lookups as the linear combinations their `VirtualPairCol`s encode (crates/stark/src/lookup/lookup.rs:10-19).
use std::collections::HashMap;
use std::rc::Rc;

use p3_air::{Air, VirtualPairCol};
use p3_field::{FieldAlgebra, PrimeField32};
use p3_koala_bear::KoalaBear;
use p3_uni_stark::{get_symbolic_constraints, Entry, SymbolicAirBuilder, SymbolicExpression};
use zkm_stark::{air::{LookupScope, MachineAir}, Chip, StarkGenericConfig, StarkMachine, PROOF_MAX_NUM_PVS};

type F = KoalaBear;

// node opcodes of the descriptor
const N_CONST: u32 = 0;
const N_MAIN: u32 = 1;
const N_PREP: u32 = 2;
const N_PUB: u32 = 3;
const N_IS_FIRST: u32 = 4;
const N_IS_LAST: u32 = 5;
const N_IS_TRANS: u32 = 6;
const N_ADD: u32 = 7;
const N_SUB: u32 = 8;
const N_MUL: u32 = 9;
const N_NEG: u32 = 10;

fn put_str(out: &mut Vec<u32>, s: &str) {
    out.push(s.len() as u32);
    for chunk in s.as_bytes().chunks(4) {
        let mut w = 0u32;
        for (i, b) in chunk.iter().enumerate() {
            w |= (*b as u32) << (8 * i);
        }
        out.push(w);
    }
}

/// Flattens the expression DAG into topologically ordered (op, a, b) triples.  Shared sub-expressions
/// (the `Rc`s of SymbolicExpression) are emitted once: the map is keyed by the Rc's address.
struct Nodes {
    triples: Vec<[u32; 3]>,
    seen: HashMap<*const SymbolicExpression<F>, u32>,
}
impl Nodes {
    fn leaf(&mut self, op: u32, a: u32, b: u32) -> u32 {
        self.triples.push([op, a, b]);
        (self.triples.len() - 1) as u32
    }
    fn rc(&mut self, e: &Rc<SymbolicExpression<F>>) -> u32 {
        let key = Rc::as_ptr(e);
        if let Some(&id) = self.seen.get(&key) {
            return id;
        }
        let id = self.expr(e);
        self.seen.insert(key, id);
        id
    }
    fn expr(&mut self, e: &SymbolicExpression<F>) -> u32 {
        match e {
            SymbolicExpression::Constant(c) => self.leaf(N_CONST, c.as_canonical_u32(), 0),
            SymbolicExpression::IsFirstRow => self.leaf(N_IS_FIRST, 0, 0),
            SymbolicExpression::IsLastRow => self.leaf(N_IS_LAST, 0, 0),
            SymbolicExpression::IsTransition => self.leaf(N_IS_TRANS, 0, 0),
            SymbolicExpression::Variable(v) => match v.entry {
                Entry::Main { offset } => self.leaf(N_MAIN, v.index as u32, offset as u32),
                Entry::Preprocessed { offset } => self.leaf(N_PREP, v.index as u32, offset as u32),
                Entry::Public => self.leaf(N_PUB, v.index as u32, 0),
                // the permutation trace and its challenges never appear in `Air::eval` of a chip: the LogUp
                // constraints are derived from the lookups on the other side of the ABI
                _ => panic!("zkm-b200: permutation/challenge variables are not part of a chip's own constraints"),
            },
            SymbolicExpression::Add { x, y, .. } => { let (a, b) = (self.rc(x), self.rc(y)); self.leaf(N_ADD, a, b) }
            SymbolicExpression::Sub { x, y, .. } => { let (a, b) = (self.rc(x), self.rc(y)); self.leaf(N_SUB, a, b) }
            SymbolicExpression::Mul { x, y, .. } => { let (a, b) = (self.rc(x), self.rc(y)); self.leaf(N_MUL, a, b) }
            SymbolicExpression::Neg { x, .. } => { let a = self.rc(x); self.leaf(N_NEG, a, 0) }
        }
    }
}

/// constant, n_terms, n_terms x (is_main, column, weight).  `VirtualPairCol` keeps its weights private, so they
/// are read back through its public `apply`: the constant is its value on all-zero rows, the weight of a
/// column its value on that column's unit vector minus the constant.
fn put_vpc(out: &mut Vec<u32>, v: &VirtualPairCol<F>, prep_w: usize, main_w: usize) {
    let zp = vec![F::ZERO; prep_w];
    let zm = vec![F::ZERO; main_w];
    let constant: F = v.apply::<F, F>(&zp, &zm);
    let mut terms = Vec::new();
    for (is_main, w) in [(0u32, prep_w), (1u32, main_w)] {
        for col in 0..w {
            let (mut p, mut m) = (zp.clone(), zm.clone());
            if is_main == 1 { m[col] = F::ONE } else { p[col] = F::ONE }
            let weight = v.apply::<F, F>(&p, &m) - constant;
            if weight != F::ZERO {
                terms.extend_from_slice(&[is_main, col as u32, weight.as_canonical_u32()]);
            }
        }
    }
    out.push(constant.as_canonical_u32());
    out.push((terms.len() / 3) as u32);
    out.extend(terms);
}

/// The descriptor of every chip of `machine` plus the FRI parameters the prover needs.
pub fn export_zkmd<SC, A>(machine: &StarkMachine<SC, A>, log_blowup: u32, num_queries: u32, pow_bits: u32) -> Vec<u32>
where
    SC: StarkGenericConfig<Val = F>,
    A: MachineAir<F> + Air<SymbolicAirBuilder<F>>,
{
    let mut out = vec![0x444d_4b5au32, 1, machine.chips().len() as u32, machine.num_pv_elts() as u32, log_blowup, num_queries, pow_bits];
    for chip in machine.chips() {
        export_chip(&mut out, chip);
    }
    out
}

fn export_chip<A>(out: &mut Vec<u32>, chip: &Chip<F, A>)
where
    A: MachineAir<F> + Air<SymbolicAirBuilder<F>>,
{
    let (prep_w, main_w) = (chip.preprocessed_width(), chip.width());
    put_str(out, &chip.name());
    out.extend_from_slice(&[
        prep_w as u32,
        main_w as u32,
        chip.log_quotient_degree() as u32,
        chip.local_only() as u32,
        (chip.commit_scope() == LookupScope::Global) as u32,
        chip.sends().len() as u32,
        chip.receives().len() as u32,
    ]);
    // constraints in assert order, as a DAG
    let constraints = get_symbolic_constraints(chip.air(), prep_w, PROOF_MAX_NUM_PVS);
    let mut nodes = Nodes { triples: Vec::new(), seen: HashMap::new() };
    let ids: Vec<u32> = constraints.iter().map(|c| nodes.expr(c)).collect();
    out.push(nodes.triples.len() as u32);
    out.push(ids.len() as u32);
    for l in chip.sends().iter().chain(chip.receives().iter()) {
        out.push(l.kind as u32);
        out.push((l.scope == LookupScope::Global) as u32);
        out.push(l.values.len() as u32);
        put_vpc(out, &l.multiplicity, prep_w, main_w);
        for v in &l.values {
            put_vpc(out, v, prep_w, main_w);
        }
    }
    for t in &nodes.triples {
        out.extend_from_slice(t);
    }
    out.extend(ids);
}

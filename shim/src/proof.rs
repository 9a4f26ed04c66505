//! Decoder of the flat "ZKPF" proof words (include/zkb200.h) into the reference's `ShardProof`
//! (crates/stark/src/types.rs:76-83) and Plonky3's `FriProof`.  Field order mirrors
//! `ziren_b200/proof.py::parse`, which the repository's tests exercise.
use hashbrown::HashMap;
use p3_commit::BatchOpening;
use p3_field::{extension::BinomialExtensionField, FieldAlgebra, FieldExtensionAlgebra};
use p3_fri::{CommitPhaseProofStep, FriProof, QueryProof};
use p3_koala_bear::KoalaBear;
use zkm_stark::{
    koala_bear_poseidon2::KoalaBearPoseidon2, septic_curve::SepticCurve, septic_digest::SepticDigest,
    septic_extension::SepticExtension, AirOpenedValues, ChipOpenedValues, ShardCommitment, ShardOpenedValues, ShardProof,
};

type F = KoalaBear;
type EF = BinomialExtensionField<F, 4>;
type SC = KoalaBearPoseidon2;

struct Reader<'a> { w: &'a [u32], i: usize }
impl<'a> Reader<'a> {
    fn u(&mut self) -> u32 { let v = self.w[self.i]; self.i += 1; v }
    fn f(&mut self) -> F { F::from_canonical_u32(self.u()) }
    fn ef(&mut self) -> EF { let c: [F; 4] = core::array::from_fn(|_| self.f()); EF::from_base_slice(&c) }
    fn digest(&mut self) -> [F; 8] { core::array::from_fn(|_| self.f()) }
    fn string(&mut self) -> String {
        let n = self.u() as usize;
        let mut bytes = Vec::with_capacity(n);
        for _ in 0..(n + 3) / 4 { bytes.extend_from_slice(&self.u().to_le_bytes()); }
        bytes.truncate(n);
        String::from_utf8(bytes).expect("chip names are UTF-8")
    }
    fn opened(&mut self, width: usize) -> AirOpenedValues<EF> {
        let local = (0..width).map(|_| self.ef()).collect();
        let next = (0..width).map(|_| self.ef()).collect();
        AirOpenedValues { local, next }
    }
}

/// `words`: the buffer `zkb200_open` returned (canonical residues).
pub fn decode_zkpf(words: &[u32]) -> ShardProof<SC> {
    let mut r = Reader { w: words, i: 0 };
    assert_eq!(r.u(), 0x4650_4b5a, "not a ZKPF proof");
    assert_eq!(r.u(), 1, "unsupported ZKPF version");
    let commitment = ShardCommitment {
        main_commit: r.digest().into(),
        permutation_commit: r.digest().into(),
        quotient_commit: r.digest().into(),
    };
    let n_chips = r.u() as usize;
    let mut chips = Vec::with_capacity(n_chips);
    let mut chip_ordering = HashMap::new();
    for index in 0..n_chips {
        let name = r.string();
        let log_degree = r.u() as usize;
        let (pw, mw, ew, nq) = (r.u() as usize, r.u() as usize, r.u() as usize, r.u() as usize);
        let preprocessed = r.opened(pw);
        let main = r.opened(mw);
        let permutation = r.opened(ew);
        let quotient = (0..nq).map(|_| (0..4).map(|_| r.ef()).collect()).collect();
        let x: [F; 7] = core::array::from_fn(|_| r.f());
        let y: [F; 7] = core::array::from_fn(|_| r.f());
        let global_cumulative_sum = SepticDigest(SepticCurve { x: SepticExtension(x), y: SepticExtension(y) });
        let local_cumulative_sum = r.ef();
        chips.push(ChipOpenedValues { preprocessed, main, permutation, quotient, global_cumulative_sum, local_cumulative_sum, log_degree });
        chip_ordering.insert(name, index);
    }
    let n_pv = r.u() as usize;
    let public_values = (0..n_pv).map(|_| r.f()).collect();
    let n_commits = r.u() as usize;
    let commit_phase_commits = (0..n_commits).map(|_| r.digest().into()).collect();
    let final_poly = r.ef();
    let pow_witness = r.f();
    let n_queries = r.u() as usize;
    let mut query_proofs = Vec::with_capacity(n_queries);
    for _ in 0..n_queries {
        let n_rounds = r.u() as usize;
        let mut input_proof = Vec::with_capacity(n_rounds);
        for _ in 0..n_rounds {
            let n_mats = r.u() as usize;
            let opened_values = (0..n_mats).map(|_| { let w = r.u() as usize; (0..w).map(|_| r.f()).collect() }).collect();
            let depth = r.u() as usize;
            let opening_proof = (0..depth).map(|_| r.digest()).collect();
            input_proof.push(BatchOpening { opened_values, opening_proof });
        }
        let n_layers = r.u() as usize;
        let mut commit_phase_openings = Vec::with_capacity(n_layers);
        for _ in 0..n_layers {
            let sibling_value = r.ef();
            let depth = r.u() as usize;
            let opening_proof = (0..depth).map(|_| r.digest()).collect();
            commit_phase_openings.push(CommitPhaseProofStep { sibling_value, opening_proof });
        }
        query_proofs.push(QueryProof { input_proof, commit_phase_openings });
    }
    assert_eq!(r.i, words.len(), "trailing words in proof");
    ShardProof {
        commitment,
        opened_values: ShardOpenedValues { chips },
        opening_proof: FriProof { commit_phase_commits, query_proofs, final_poly, pow_witness },
        chip_ordering,
        public_values,
    }
}

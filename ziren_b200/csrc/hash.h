// K2: Poseidon2 Merkle-tree (MMCS) build, batched permutation, proof-of-work grinding, and the
// host-side duplex challenger.  Replaces [P3-upstream] MerkleTreeMmcs::commit / PaddingFreeSponge /
// TruncatedPermutation / DuplexChallenger as reached from TwoAdicFriPcs::commit and the FRI commit
// phase (reference call sites crates/stark/src/prover.rs:277,403,497,547; structure pinned by
// crates/recursion/circuit/src/fri.rs:363-405, hash.rs:40-80, challenger.rs:60-233).
#pragma once
#include <map>
#include "common.h"
#include "poseidon2.cuh"

namespace zkb {

struct MatRef {
  const u32* ptr;   // column-major
  u32 width;
  u32 log_height;
};

void p2_upload_constants();  // once per device

// Digest layers of one commitment, each SoA: word j of node i at layer[j * count + i].
struct DigestLayers {
  DevBuf buf;
  std::vector<size_t> offset;  // per layer, in words
  std::vector<size_t> count;   // nodes per layer; layer 0 = leaves
  u32* layer(size_t l) const { return buf.p + offset[l]; }
};

// Build the tree over column-major matrices of power-of-two heights (leaf = sponge over the
// concatenated rows of the tallest matrices in list order; shorter ones injected on the way up).
// `mats_dev` is the same list in device memory.  Root (Montgomery) is written to root_dev[8].
// `pre`: row digests ([8][height] word-major) already computed for whole height groups, keyed by log
// height (leaf_absorb below); those groups are not hashed again.
void merkle_build(const std::vector<MatRef>& mats, ParamArena& arena, DigestLayers& out, u32* root_dev, cudaStream_t s,
                  const std::map<unsigned, const u32*>* pre = nullptr);
// One piece (a device array of column pointers, all of `height` rows) of a height class into the per-row
// sponge states (state: [16][height] words); see hash.cu.
void leaf_absorb(const u32* const* cols_dev, size_t height, u32 ncols, u32* state, bool first, bool last, u32* digests, cudaStream_t s);

// Tree over one FRI commit-phase layer: folded = m EF values, component-major [4][m].
void fri_commit_layer(const u32* folded, size_t m, DigestLayers& out, u32* root_dev, cudaStream_t s);

void permute_batch(u32* states, size_t n, cudaStream_t s);   // n x 16 row-major Montgomery states

// Smallest witness w (canonical) with sample_bits(bits) == 0 after observing w.
// st: sponge state (Montgomery) with the pending inputs already written to st[0..n_in).
u32 grind_witness(const u32 st[16], unsigned n_in, unsigned bits, u32* scratch_dev, cudaStream_t s);

// The same duplex challenger RESIDENT ON THE DEVICE for the FRI commit phase ([P3-upstream] fri::prover::commit_phase:
// observe the layer's commitment, sample beta, fold; structure pinned by crates/recursion/circuit/src/fri.rs:247-361):
// the transcript of the ~20 layers advances in one-thread kernels between the tree and fold kernels, so the host
// enqueues the whole phase without a round trip per layer and reads roots, final polynomial and challenger back once.
struct DevChallenger { u32 state[16]; u32 in_buf[8]; u32 out_buf[8]; u32 n_in, n_out; };     // Montgomery
// dst <- the host challenger's image (passed by value: no host-to-device copy on a compute lane)
void challenger_to_device(const struct Challenger& ch, DevChallenger* dst, cudaStream_t s);
// observe the 8-word digest at root_dev (Montgomery), then sample an extension element into beta_dev[4]
void challenger_observe_digest_sample_ext(DevChallenger* ch, const u32* root_dev, u32* beta_dev, cudaStream_t s);

// DuplexChallenger<KoalaBear, Poseidon2, 16, 8> on the host (Montgomery residues internally).
struct Challenger {
  Fp state[16];
  Fp in_buf[8];
  Fp out_buf[8];
  unsigned n_in = 0, n_out = 0;
  Challenger() { for (auto& x : state) x = fp_zero(); }
  void duplexing();
  void observe(Fp v);
  void observe_canonical(u32 v) { observe(fp_from_canonical(v)); }
  void observe_digest(const u32* d_monty) { for (int i = 0; i < 8; i++) observe(fp_raw(d_monty[i])); }
  void observe_ext(const Ef& e) { for (int i = 0; i < 4; i++) observe(e.c[i]); }
  Fp sample();
  Ef sample_ext() { Ef e; for (int i = 0; i < 4; i++) e.c[i] = sample(); return e; }
  u32 sample_bits(unsigned bits) { return fp_to_canonical(sample()) & ((1u << bits) - 1); }
  // 34-word canonical image: state[16], n_in, in[8], n_out, out[8]
  void load(const u32* w);
  void store(u32* w) const;
  void load_device_image(const DevChallenger& d);     // after the FRI commit phase ran on the device
};

}  // namespace zkb

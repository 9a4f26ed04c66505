// K2: Poseidon2 sponge / compression kernels and the Merkle (MMCS) tree build.
// One thread owns one row (leaf) or one tree node: the 16-word state lives in registers,
// column-major matrices make every load of a warp one contiguous 128 B segment, digest layers
// are stored word-major (SoA) so that the two children of a node are one 8-byte load.
#include "hash.h"
#include <algorithm>
#include <map>

namespace zkb {

const P2Consts& p2_host_consts() {
  static const P2Consts c = p2_make_consts();
  return c;
}

__constant__ P2Consts d_p2;

void p2_upload_constants() {
  ZKB_CUDA(cudaMemcpyToSymbol(d_p2, &p2_host_consts(), sizeof(P2Consts)));
}

__device__ __forceinline__ void p2_permute_dev(Fp* s) { p2_permute_with(s, d_p2); }

// ---- leaf hashing: PaddingFreeSponge over the concatenated rows of same-height matrices ----
__device__ __forceinline__ void sponge_rows(const MatRef* __restrict__ mats, int nmats, size_t height, size_t r, Fp* st) {
#pragma unroll
  for (int i = 0; i < 16; i++) st[i] = fp_zero();
  int mi = 0;
  u32 col = 0;
  while (mi < nmats && mats[mi].width == 0) mi++;   // skip empty matrices
  while (mi < nmats) {
    // whole 8-column chunks inside the current matrix: eight strided loads and a permutation
    const u32 w = mats[mi].width;
    const u32* __restrict__ p = mats[mi].ptr + (size_t)col * height + r;
    while (col + 8 <= w) {
#pragma unroll
      for (int i = 0; i < 8; i++) st[i] = fp_raw(p[(size_t)i * height]);
      p += 8 * height;
      col += 8;
      p2_permute_dev(st);
    }
    if (col == w) {
      col = 0;
      do { mi++; } while (mi < nmats && mats[mi].width == 0);
      continue;
    }
    // ragged chunk: the rate block spans the end of this matrix (and maybe whole narrow ones)
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (mi < nmats) {
        st[i] = fp_raw(mats[mi].ptr[(size_t)col * height + r]);
        col++;
        while (mi < nmats && col >= mats[mi].width) { mi++; col = 0; }
      }
    }
    p2_permute_dev(st);
  }
}

__global__ void __launch_bounds__(128, 16) leaf_hash_kernel(const MatRef* __restrict__ mats, int nmats, size_t height,
                                                        u32* __restrict__ out) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= height) return;
  Fp st[16];
  sponge_rows(mats, nmats, height, r, st);
#pragma unroll
  for (int j = 0; j < 8; j++) out[j * height + r] = st[j].v;
}

static size_t leaf_smem_for(size_t nblocks);

// Piecewise leaf hashing: the sponge over the concatenated rows of one height class (its matrices in
// commit order) is absorbed in column pieces as their LDEs are produced (the upload of the next piece
// overlaps, prover_commit), the 16-word state of every row parked in HBM between pieces (SoA
// [16][height]).  A piece is a list of column pointers, so it may straddle matrices; every piece but
// the last is a whole number of rate blocks (8 columns), the last one takes the ragged tail
// (PaddingFreeSponge: a short block overwrites state[0..len) only) and writes the digests.
__global__ void __launch_bounds__(128, 16) leaf_absorb_kernel(const u32* const* __restrict__ cols, size_t height, u32 ncols,
                                                          u32* __restrict__ state, int first, int last,
                                                          u32* __restrict__ digests) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= height) return;
  Fp st[16];
  if (first) {
#pragma unroll
    for (int i = 0; i < 16; i++) st[i] = fp_zero();
  } else {
#pragma unroll
    for (int i = 0; i < 16; i++) st[i] = fp_raw(state[(size_t)i * height + r]);
  }
  u32 c = 0;
  for (; c + 8 <= ncols; c += 8) {
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = fp_raw(cols[c + i][r]);      // the column pointers are warp-uniform
    p2_permute_dev(st);
  }
  if (c < ncols) {   // ragged tail (last piece of the height class only)
#pragma unroll
    for (int i = 0; i < 8; i++) if (c + i < ncols) st[i] = fp_raw(cols[c + i][r]);
    p2_permute_dev(st);
  }
  if (last) {
#pragma unroll
    for (int j = 0; j < 8; j++) digests[j * height + r] = st[j].v;
  } else {
#pragma unroll
    for (int i = 0; i < 16; i++) state[(size_t)i * height + r] = st[i].v;
  }
}
void leaf_absorb(const u32* const* cols_dev, size_t height, u32 ncols, u32* state, bool first, bool last, u32* digests, cudaStream_t s) {
  if (!last && (ncols & 7)) throw std::runtime_error("zkb200: leaf_absorb: a non-final piece must be whole rate blocks");
  if (!ncols && !last) return;
  const unsigned nblocks = ceil_div(height, 128);
  leaf_absorb_kernel<<<nblocks, 128, leaf_smem_for(nblocks), s>>>(cols_dev, height, ncols, state, first ? 1 : 0, last ? 1 : 0, digests);
  ZKB_CHECK_LAUNCH();
}

// next[i] = compress(prev[2i], prev[2i+1]); optionally then compress(., inject[i]) where inject holds
// the row digests of the matrices of this height (hashed by leaf_hash_kernel, word-major)
__global__ void __launch_bounds__(128) compress_kernel(const u32* __restrict__ prev, u32* __restrict__ next, size_t m,
                                                       const u32* __restrict__ inject) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= m) return;
  Fp st[16];
#pragma unroll
  for (int j = 0; j < 8; j++) {
    uint2 pr = reinterpret_cast<const uint2*>(prev + j * 2 * m)[i];
    st[j] = fp_raw(pr.x);
    st[8 + j] = fp_raw(pr.y);
  }
  p2_permute_dev(st);
  if (inject) {
#pragma unroll
    for (int j = 0; j < 8; j++) st[8 + j] = fp_raw(inject[j * m + i]);
    p2_permute_dev(st);
  }
#pragma unroll
  for (int j = 0; j < 8; j++) next[j * m + i] = st[j].v;
}

// The row-hashing kernel runs one long-lived thread per row (hundreds of permutations), so a
// partially filled last wave costs real time: 2^19 rows are 4096 CTAs = 1.73 waves at 16 CTAs/SM
// but 1.98 waves at 14.  Pick the residency (12..16 CTAs per SM, enforced through a dynamic
// shared-memory reservation) whose last wave is fullest.
static size_t leaf_smem_for(size_t nblocks) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int best_b = 16;
  double best_eff = 0;
  for (int b = 16; b >= 12; b--) {
    double waves = (double)nblocks / ((double)sms * b);
    double eff = waves / (double)(size_t)(waves + 0.999999);
    if (waves <= 1.0) eff = 1.0;
    if (eff > best_eff + 0.02) { best_eff = eff; best_b = b; }
  }
  if (best_b == 16) return 0;
  return ((size_t)227 * 1024 / best_b - 1024) & ~(size_t)1023;   // leaves room for exactly best_b CTAs
}
static void launch_leaf_hash(const MatRef* mats_dev, int nmats, size_t height, u32* out, cudaStream_t s) {
  const unsigned nblocks = ceil_div(height, 128);
  leaf_hash_kernel<<<nblocks, 128, leaf_smem_for(nblocks), s>>>(mats_dev, nmats, height, out);
  ZKB_CHECK_LAUNCH();
}

// The top of a tree (<= 512 nodes wide, no injections) in one CTA: levels separated by
// __syncthreads instead of kernel launches.  layers are consecutive SoA blocks in `buf`.
__global__ void __launch_bounds__(256) compress_top_kernel(u32* buf, size_t off_first, size_t m_first, int nlevels) {
  size_t off_prev = off_first, m = m_first;   // prev layer has 2m nodes
  for (int l = 0; l < nlevels; l++) {
    const u32* prev = buf + off_prev;
    u32* next = buf + off_prev + 16 * m;
    for (size_t i = threadIdx.x; i < m; i += blockDim.x) {
      Fp st[16];
#pragma unroll
      for (int j = 0; j < 8; j++) {
        st[j] = fp_raw(prev[j * 2 * m + 2 * i]);
        st[8 + j] = fp_raw(prev[j * 2 * m + 2 * i + 1]);
      }
      p2_permute_dev(st);
#pragma unroll
      for (int j = 0; j < 8; j++) next[j * m + i] = st[j].v;
    }
    __syncthreads();
    off_prev += 16 * m;
    m >>= 1;
  }
}

static void alloc_layers(DigestLayers& out, unsigned max_log, cudaStream_t s) {
  out.offset.clear(); out.count.clear();
  size_t total = 0;
  for (unsigned l = 0; l <= max_log; l++) {
    size_t cnt = (size_t)1 << (max_log - l);
    out.offset.push_back(total);
    out.count.push_back(cnt);
    total += 8 * cnt;
  }
  out.buf = DevBuf(total, s);
}

// layers 1..max_log from the leaf layer; `groups` = matrices to inject, keyed by log height
static void build_upper(DigestLayers& out, unsigned max_log, const std::map<unsigned, std::vector<MatRef>>& groups,
                        std::map<unsigned, const MatRef*>& dev_groups, u32* root_dev, cudaStream_t s,
                        const std::map<unsigned, const u32*>* pre = nullptr) {
  unsigned l = 1;
  while (l <= max_log) {
    size_t m = out.count[l];
    unsigned lh = max_log - l;
    bool inject = groups.count(lh) != 0;
    if (!inject && m <= 512) {
      // run as many injection-free levels as possible inside one CTA
      unsigned l2 = l;
      while (l2 <= max_log && groups.count(max_log - l2) == 0) l2++;
      compress_top_kernel<<<1, 256, 0, s>>>(out.buf.p, out.offset[l - 1], m, (int)(l2 - l));
      ZKB_CHECK_LAUNCH();
      l = l2;
      continue;
    }
    const u32* pre_inj = (inject && pre && pre->count(lh)) ? pre->at(lh) : nullptr;
    DevBuf inj(inject && !pre_inj ? 8 * m : 0, s);
    if (inject && !pre_inj) launch_leaf_hash(dev_groups[lh], (int)groups.at(lh).size(), m, inj.p, s);
    compress_kernel<<<ceil_div(m, 128), 128, 0, s>>>(out.layer(l - 1), out.layer(l), m, inject ? (pre_inj ? pre_inj : inj.p) : nullptr);
    ZKB_CHECK_LAUNCH();
    l++;
  }
  ZKB_CUDA(cudaMemcpyAsync(root_dev, out.layer(max_log), 8 * sizeof(u32), cudaMemcpyDeviceToDevice, s));
}

void merkle_build(const std::vector<MatRef>& mats, ParamArena& arena, DigestLayers& out, u32* root_dev, cudaStream_t s,
                  const std::map<unsigned, const u32*>* pre) {
  unsigned max_log = 0;
  for (auto& m : mats) max_log = std::max(max_log, m.log_height);
  // group by height, list order preserved
  std::map<unsigned, std::vector<MatRef>> groups;
  for (auto& m : mats) groups[m.log_height].push_back(m);
  std::map<unsigned, const MatRef*> dev_groups;
  for (auto& g : groups) dev_groups[g.first] = arena.push(g.second.data(), g.second.size());
  alloc_layers(out, max_log, s);
  size_t h = (size_t)1 << max_log;
  if (pre && pre->count(max_log)) ZKB_CUDA(cudaMemcpyAsync(out.layer(0), pre->at(max_log), 8 * h * sizeof(u32), cudaMemcpyDeviceToDevice, s));
  else launch_leaf_hash(dev_groups[max_log], (int)groups[max_log].size(), h, out.layer(0), s);
  std::map<unsigned, std::vector<MatRef>> inj = groups;
  inj.erase(max_log);
  build_upper(out, max_log, inj, dev_groups, root_dev, s, pre);
}

// FRI commit-phase layer: leaves are the pairs (e_{2i}, e_{2i+1}) of the folded vector flattened
// to 8 base elements (ExtensionMmcs over a width-2 EF matrix) -> exactly one permutation per leaf.
__global__ void __launch_bounds__(128) fri_leaf_kernel(const u32* __restrict__ folded, size_t m, u32* __restrict__ out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t half = m >> 1;
  if (i >= half) return;
  Fp st[16];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    uint2 pr = reinterpret_cast<const uint2*>(folded + (size_t)c * m)[i];
    st[c] = fp_raw(pr.x);
    st[4 + c] = fp_raw(pr.y);
  }
#pragma unroll
  for (int j = 8; j < 16; j++) st[j] = fp_zero();
  p2_permute_dev(st);
#pragma unroll
  for (int j = 0; j < 8; j++) out[j * half + i] = st[j].v;
}
void fri_commit_layer(const u32* folded, size_t m, DigestLayers& out, u32* root_dev, cudaStream_t s) {
  const size_t half = m >> 1;
  unsigned max_log = log2_exact(half);
  alloc_layers(out, max_log, s);
  fri_leaf_kernel<<<ceil_div(half, 128), 128, 0, s>>>(folded, m, out.layer(0));
  ZKB_CHECK_LAUNCH();
  std::map<unsigned, std::vector<MatRef>> none;
  std::map<unsigned, const MatRef*> none_dev;
  build_upper(out, max_log, none, none_dev, root_dev, s);
}

__global__ void permute_batch_kernel(u32* states, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp st[16];
#pragma unroll
  for (int j = 0; j < 16; j++) st[j] = fp_raw(states[16 * i + j]);
  p2_permute_dev(st);
#pragma unroll
  for (int j = 0; j < 16; j++) states[16 * i + j] = st[j].v;
}
void permute_batch(u32* states, size_t n, cudaStream_t s) {
  if (!n) return;
  permute_batch_kernel<<<ceil_div(n, 128), 128, 0, s>>>(states, n);
  ZKB_CHECK_LAUNCH();
}

// ---- proof-of-work grinding ---------------------------------------------------------------------
struct GrindArgs { u32 st[16]; };
__global__ void grind_kernel(GrindArgs a, unsigned n_in, u32 mask, u32 base, u32* result) {
  u32 w = base + blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= KB_P) return;
  Fp st[16];
#pragma unroll
  for (int j = 0; j < 16; j++) st[j] = fp_raw(a.st[j]);
  Fp wm = fp_from_canonical(w);
#pragma unroll
  for (int j = 0; j < 8; j++) if (j == (int)n_in) st[j] = wm;
  p2_permute_dev(st);
  if ((fp_to_canonical(st[7]) & mask) == 0) atomicMin(result, w);
}
u32 grind_witness(const u32 st[16], unsigned n_in, unsigned bits, u32* scratch_dev, cudaStream_t s) {
  GrindArgs a;
  memcpy(a.st, st, sizeof(a.st));
  const u32 mask = (1u << bits) - 1;
  const u32 batch = 1u << 20;
  u32 init = 0xffffffffu, found = 0xffffffffu;
  (void)init;
  ZKB_CUDA(cudaMemsetAsync(scratch_dev, 0xff, 4, s));      // no host->device copy on a compute lane (common.h: pull_words)
  for (u64 base = 0; base < KB_P; base += batch) {
    grind_kernel<<<batch / 256, 256, 0, s>>>(a, n_in, mask, (u32)base, scratch_dev);
    ZKB_CHECK_LAUNCH();
    ZKB_CUDA(cudaMemcpyAsync(&found, scratch_dev, 4, cudaMemcpyDeviceToHost, s));
    ZKB_CUDA(cudaStreamSynchronize(s));
    if (found != 0xffffffffu) return found;
  }
  throw std::runtime_error("zkb200: proof-of-work witness not found");
}

// ---- device-resident challenger (FRI commit phase) ----------------------------------------------------
__device__ void dch_duplexing(DevChallenger& c) {
  Fp st[16];
  for (int i = 0; i < 16; i++) st[i] = fp_raw(i < (int)c.n_in ? c.in_buf[i & 7] : c.state[i]);
  c.n_in = 0;
  p2_permute_dev(st);
  for (int i = 0; i < 16; i++) c.state[i] = st[i].v;
  for (int i = 0; i < 8; i++) c.out_buf[i] = st[i].v;
  c.n_out = 8;
}
__device__ void dch_observe(DevChallenger& c, u32 v) {
  c.n_out = 0;
  c.in_buf[c.n_in++] = v;
  if (c.n_in == 8) dch_duplexing(c);
}
__device__ u32 dch_sample(DevChallenger& c) {
  if (c.n_in != 0 || c.n_out == 0) dch_duplexing(c);
  return c.out_buf[--c.n_out];
}
__global__ void challenger_set_kernel(DevChallenger* dst, DevChallenger v) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *dst = v;
}
__global__ void challenger_observe_sample_kernel(DevChallenger* ch, const u32* __restrict__ root, u32* __restrict__ beta) {
  if (blockIdx.x || threadIdx.x) return;
  DevChallenger c = *ch;
  for (int i = 0; i < 8; i++) dch_observe(c, root[i]);
  for (int i = 0; i < 4; i++) beta[i] = dch_sample(c);
  *ch = c;
}
void challenger_to_device(const Challenger& ch, DevChallenger* dst, cudaStream_t s) {
  DevChallenger v;
  for (int i = 0; i < 16; i++) v.state[i] = ch.state[i].v;
  for (int i = 0; i < 8; i++) { v.in_buf[i] = i < (int)ch.n_in ? ch.in_buf[i].v : 0; v.out_buf[i] = i < (int)ch.n_out ? ch.out_buf[i].v : 0; }
  v.n_in = ch.n_in; v.n_out = ch.n_out;
  challenger_set_kernel<<<1, 32, 0, s>>>(dst, v);
  ZKB_CHECK_LAUNCH();
}
void challenger_observe_digest_sample_ext(DevChallenger* ch, const u32* root_dev, u32* beta_dev, cudaStream_t s) {
  challenger_observe_sample_kernel<<<1, 32, 0, s>>>(ch, root_dev, beta_dev);
  ZKB_CHECK_LAUNCH();
}
void Challenger::load_device_image(const DevChallenger& d) {
  if (d.n_in >= 8 || d.n_out > 8) throw std::runtime_error("zkb200: malformed device challenger state");
  for (int i = 0; i < 16; i++) state[i] = fp_raw(d.state[i]);
  n_in = d.n_in; n_out = d.n_out;
  for (int i = 0; i < 8; i++) { in_buf[i] = fp_raw(d.in_buf[i]); out_buf[i] = fp_raw(d.out_buf[i]); }
}

// ---- host challenger ------------------------------------------------------------------------------
void Challenger::duplexing() {
  for (unsigned i = 0; i < n_in; i++) state[i] = in_buf[i];
  n_in = 0;
  p2_permute_host(state);
  for (int i = 0; i < 8; i++) out_buf[i] = state[i];
  n_out = 8;
}
void Challenger::observe(Fp v) {
  n_out = 0;
  in_buf[n_in++] = v;
  if (n_in == 8) duplexing();
}
Fp Challenger::sample() {
  if (n_in != 0 || n_out == 0) duplexing();
  return out_buf[--n_out];
}
void Challenger::load(const u32* w) {
  for (int i = 0; i < 16; i++) state[i] = fp_from_canonical(w[i]);
  n_in = w[16];
  for (int i = 0; i < 8; i++) in_buf[i] = fp_from_canonical(w[17 + i]);
  n_out = w[25];
  for (int i = 0; i < 8; i++) out_buf[i] = fp_from_canonical(w[26 + i]);
  // a duplex challenger never rests with a full input buffer (observe() permutes at 8)
  if (n_in >= 8 || n_out > 8) throw std::runtime_error("zkb200: malformed challenger state");
}
void Challenger::store(u32* w) const {
  for (int i = 0; i < 16; i++) w[i] = fp_to_canonical(state[i]);
  w[16] = n_in;
  for (int i = 0; i < 8; i++) w[17 + i] = i < (int)n_in ? fp_to_canonical(in_buf[i]) : 0;
  w[25] = n_out;
  for (int i = 0; i < 8; i++) w[26 + i] = i < (int)n_out ? fp_to_canonical(out_buf[i]) : 0;
}

}  // namespace zkb

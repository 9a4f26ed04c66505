// K1: batched KoalaBear NTT / coset LDE over column-major matrices.
//
// A transform of length n = 2^(K1+K2) is TWO levels (four-step): a STRIDED level of 2^K1-point
// NTTs over elements 2^K2 apart, and a CONTIGUOUS level of 2^K2-point NTTs, linked by the twiddle
// w_n^(k1*t).  Each level runs inside shared-memory tiles; inside a tile every thread keeps 16
// elements in registers and performs radix-2..16 butterfly rounds there (<= 3 shared-memory round
// trips for 2^12 points).  The coset LDE is three launches per column chunk and never needs a
// separate bit-reversal, transpose or scaling pass:
//   A  strided DIF (inverse twiddles)                       evaluations      -> half-transformed
//   B  contiguous DIF  ->  x shift_c^i / n  ->  contiguous DIT, once per coset c (fused in smem)
//   C  strided DIT with the BIT-REVERSED store of the committed row order
// (n <= 2^12: B alone, storing bit-reversed).  The first radix round of a level reads straight from
// global memory and the last one writes straight to it.  The kernels are bound by integer
// instruction issue, not by HBM (profiles/README.md), so columns are processed in large chunks
// (2^26 elements) that keep every launch several waves deep rather than L2-sized ones.
#include "ntt.h"
#include <algorithm>
#include <atomic>
#include <cstdlib>

namespace zkb {

static constexpr int TW_SPLIT = 12;
static constexpr int KMAX = 12;             // largest in-tile transform
static constexpr int ELEMS_PER_THREAD = 16;

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
// run-time experiment knobs (zkb200_set_option): 0 = built-in choice
std::atomic<int> g_ntt_force_k2{0};

__global__ void tw_init_kernel(u32* lo, u32* hi) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1 << TW_SPLIT)) return;
  Fp w = two_adic_generator(24);
  lo[i] = fp_pow(w, (u64)i).v;
  hi[i] = fp_pow(w, (u64)i << TW_SPLIT).v;
}
// Shoup pair of a Montgomery-form constant c: (w, w') with w = canonical(c), w' = floor(w 2^32 / p).
// For any 32-bit a:  a*w - umulhi(a, w')*p  lies in [0, 2p) and is = a*w (mod p); with `a` a
// Montgomery residue the product is again a Montgomery residue.
__device__ __forceinline__ uint2 shoup_pair(Fp c) {
  u32 w = fp_to_canonical(c);
  return make_uint2(w, (u32)((((u64)w) << 32) / KB_P));
}
__device__ __forceinline__ Fp shoup_mul(u32 a, uint2 w) {
  u32 q = __umulhi(a, w.y);
  u32 r = a * w.x - q * KB_P;
  u32 t = r - KB_P;
  return fp_raw(t < r ? t : r);
}

// small[dir][K] : w_{2^K}^(+-e), e < 2^(K-1), as Shoup pairs (K <= KMAX_TAB)
static constexpr int KMAX_TAB = 12;
__global__ void small_tw_kernel(uint2* out, const u32* lo, const u32* hi) {
  // layout: for dir in {fwd, inv}: for K in 1..12: 2^(K-1) entries at offset (2^(K-1) - 1)
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  const u32 per_dir = (1u << KMAX_TAB) - 1;
  if (i >= 2 * per_dir) return;
  u32 dir = i / per_dir, r = i % per_dir;
  int K = 32 - __clz(r + 1);            // r in [2^(K-1)-1, 2^K-1)
  u32 e = r - ((1u << (K - 1)) - 1);
  u32 E = e << (24 - K);
  if (dir) E = (0u - E) & ((1u << 24) - 1);
  out[i] = shoup_pair(tw_pow2(lo, hi, E));
}
void ntt_set_device_attributes();
void NttTables::init(cudaStream_t s) {
  ntt_set_device_attributes();
  ZKB_CUDA(cudaMalloc((void**)&tw_lo, sizeof(u32) << TW_SPLIT));
  ZKB_CUDA(cudaMalloc((void**)&tw_hi, sizeof(u32) << TW_SPLIT));
  tw_init_kernel<<<(1 << TW_SPLIT) / 256, 256, 0, s>>>(tw_lo, tw_hi);
  ZKB_CHECK_LAUNCH();
  const u32 cnt = 2 * ((1u << KMAX_TAB) - 1);
  ZKB_CUDA(cudaMalloc((void**)&small_tw, cnt * sizeof(uint2)));
  small_tw_kernel<<<ceil_div(cnt, 256), 256, 0, s>>>((uint2*)small_tw, tw_lo, tw_hi);
  ZKB_CHECK_LAUNCH();
}
void NttTables::destroy() {
  if (tw_lo) cudaFree(tw_lo);
  if (tw_hi) cudaFree(tw_hi);
  if (small_tw) cudaFree(small_tw);
  for (auto& kv : four_step) cudaFree(kv.second);
  four_step.clear(); scale_cache.clear(); scale_bytes.clear(); scale_cache_bytes = 0;
  tw_lo = tw_hi = nullptr; small_tw = nullptr;
}

// four-step twiddle of the element stored at position o*S + t of a 2^(K1+K2) column:
// w_n^(+-bitrev_K1(o) * t), as a Shoup pair, indexed by the position itself (coalesced, L2-resident)
__global__ void four_step_kernel(uint2* out, int K1, int logS, int inverse, const u32* lo, const u32* hi) {
  const u32 logn = K1 + logS;
  size_t pos = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (pos >= ((size_t)1 << logn)) return;
  u32 o = (u32)(pos >> logS), t = (u32)pos & ((1u << logS) - 1);
  u32 e = (bitrev32(o, K1) * t) & ((1u << logn) - 1);
  u32 E = e << (24 - logn);
  if (inverse) E = (0u - E) & ((1u << 24) - 1);
  out[pos] = shoup_pair(tw_pow2(lo, hi, E));
}
const void* NttTables::four_step_table(int K1, int logS, bool inverse, cudaStream_t s) const {
  std::lock_guard<std::mutex> lock(mu);
  const int key = ((K1 + logS) << 1) | (inverse ? 1 : 0);
  auto it = four_step.find(key);
  if (it != four_step.end()) return it->second;
  const size_t n = (size_t)1 << (K1 + logS);
  void* p = nullptr;
  ZKB_CUDA(cudaMalloc(&p, n * sizeof(uint2)));
  four_step_kernel<<<ceil_div(n, 256), 256, 0, s>>>((uint2*)p, K1, logS, inverse ? 1 : 0, tw_lo, tw_hi);
  ZKB_CHECK_LAUNCH();
  ZKB_CUDA(cudaStreamSynchronize(s));   // published to every lane: must be complete
  four_step[key] = p;
  return p;
}

// threads of a CTA whose tile holds 2^(K+LT) elements, 16 per thread (0: not known at compile time)
__host__ __device__ constexpr int tile_threads(int K, int LT) { return LT < 0 ? 0 : (K + LT - 4 >= 5 ? 1 << (K + LT - 4) : 32); }

// ---- in-register butterfly rounds --------------------------------------------------------------
// Everything about a round is a compile-time constant (transform size 2^K, radix 2^R, bit offset
// LO), so twiddle indices and element offsets fold into immediates.
// x[m] holds slot (base | m << LO) of a 2^K-point transform; bl = base & (2^LO - 1).
// tw[e] = Shoup pair of w_{2^K}^(+-e), e < 2^(K-1).
template <int K, int R, int LO, bool DIF>
__device__ __forceinline__ void butterflies(Fp* x, u32 bl, const uint2* __restrict__ tw) {
#pragma unroll
  for (int j = 0; j < R; j++) {
    const int lh = DIF ? (R - 1 - j) : j;   // log2(half) in m-space
    const int sh = K - 1 - LO - lh;         // slot-space half is 2^(LO+lh)
    const u32 tbase = bl << sh;
#pragma unroll
    for (int p = 0; p < (1 << (R - 1)); p++) {
      const int k = p & ((1 << lh) - 1);
      const int m0 = ((p >> lh) << (lh + 1)) | k;
      const int m1 = m0 + (1 << lh);
      const u32 e = tbase + ((u32)k << (LO + sh));
      const uint2 w = tw[e + (e >> 4)];          // padded: strided twiddle reads stay conflict-free
      if (DIF) {
        Fp a = x[m0], b = x[m1];
        x[m0] = a + b;
        x[m1] = shoup_mul(a.v - b.v + KB_P, w);     // a - b + p in (0, 2p): any u32 is fine
      } else {
        Fp t = shoup_mul(x[m1].v, w);
        x[m1] = x[m0] - t;
        x[m0] = x[m0] + t;
      }
    }
  }
}

// Shared-memory layouts.  Strided tiles: row = slot, 2^logT offsets per row, XOR-swizzled so that
// both "lanes along q" and "lanes along slot" accesses are conflict-free without padding:
//   saddr(slot, q) = slot * T + (q ^ (slot & (T-1)))
// Contiguous tiles: one row per group, addr = q * ldg + slot + slot/8 (lanes run along slot).
__device__ __forceinline__ u32 saddr(u32 slot, u32 q, int logT) { return (slot << logT) + (q ^ (slot & ((1u << logT) - 1))); }
__device__ __forceinline__ u32 caddr(u32 slot, u32 q, u32 ldg) { return q * ldg + slot + (slot >> 3); }

// One radix-2^R round over bits [LO, LO+R) of the slot index, 16 elements per thread.  Where the
// elements come from and go to is given by the caller (shared memory or straight from / to global
// memory): the first round of a level loads from global memory into registers and the last one
// stores from registers, so a 2^9-point level makes one or two shared-memory round trips.
// NT > 0: the CTA size is known at compile time (tile width fixed by the template), so the group
// guards and index splits fold away; NT == 0 reads blockDim.
template <int K, int R, int LO, bool DIF, bool CONTIG, int NT, class Load, class Store>
__device__ __forceinline__ void round_io(int logT, const uint2* __restrict__ tw, Load load, Store store) {
  constexpr int G = ELEMS_PER_THREAD >> R;   // groups per thread
  const u32 nthreads = NT > 0 ? (u32)NT : blockDim.x;
  const u32 T = 1u << logT;
  const u32 ngroups = 1u << (K - R + logT);
  Fp x[G][1 << R];
  u32 qs[G], bases[G];
#pragma unroll
  for (int g = 0; g < G; g++) {
    const u32 gi = g * nthreads + threadIdx.x;
    u32 rest, q;
    if (CONTIG) { rest = gi & ((1u << (K - R)) - 1); q = gi >> (K - R); }
    else { q = gi & (T - 1); rest = gi >> logT; }
    qs[g] = q;
    bases[g] = ((rest >> LO) << (LO + R)) | (rest & ((1u << LO) - 1));
    if (gi < ngroups) {
#pragma unroll
      for (int m = 0; m < (1 << R); m++) x[g][m] = load(bases[g] | ((u32)m << LO), q);
    }
  }
#pragma unroll
  for (int g = 0; g < G; g++) {
    const u32 gi = g * nthreads + threadIdx.x;
    if (gi < ngroups) {
      butterflies<K, R, LO, DIF>(x[g], bases[g] & ((1u << LO) - 1), tw);
#pragma unroll
      for (int m = 0; m < (1 << R); m++) store(bases[g] | ((u32)m << LO), qs[g], x[g][m]);
    }
  }
}

// rounds of radix <= 16, evenly split: R_r = K/nr + (r < K%nr)
__host__ __device__ constexpr int round_bits(int K, int r) { return K / ((K + 3) / 4) + (r < K % ((K + 3) / 4) ? 1 : 0); }
// bit offset of round r when the rounds are laid out from the top bit down (r = 0 on top)
__host__ __device__ constexpr int round_lo(int K, int r) {
  int lo = K;
  for (int i = 0; i <= r; i++) lo -= round_bits(K, i);
  return lo;
}

template <int K>
__device__ __forceinline__ void fill_small_twiddles(uint2* tw, bool inverse, const uint2* __restrict__ small_tw) {
  constexpr u32 half = 1u << (K - 1);
  const uint2* src = small_tw + (inverse ? ((1u << KMAX_TAB) - 1) : 0) + (half - 1);
  for (u32 m = threadIdx.x; m < half; m += blockDim.x) tw[m + (m >> 4)] = src[m];
}
// words of shared memory taken by one padded twiddle table of a 2^K-point transform
__host__ __device__ constexpr u32 tw_words(int K) { return ((1u << K) + ((1u << K) >> 4) + 4) & ~3u; }

// ---- strided level -----------------------------------------------------------------------------
struct StridedArgs {
  const u32* in; u32* out;
  size_t in_stride, out_stride;      // elements between columns
  const uint2* small_tw;
  const uint2* four;                 // four-step twiddles indexed by element position
  int logS, logT;                    // n = 2^(K+logS)
  int inverse;
  int final_dit;                     // 0: DIF with post-twiddle, in-place positions
                                     // 1: DIT (pre-twiddle) storing bit-reversed positions
  size_t out_coset_stride;           // final_dit: blockIdx.z selects the coset block of in/out
  size_t in_coset_stride;
};

// LT >= 0: tile width 2^LT known at compile time (the launcher's default for this K), which folds the
// shared-memory address arithmetic into immediates (13-22 % fewer instructions); LT < 0: read it
// from the arguments (small transforms whose tiles are clipped)
template <int K, bool FINAL, int LT>
__device__ __forceinline__ void strided_body(const StridedArgs& a, u32* smem) {
  const int logT = LT >= 0 ? LT : a.logT, logS = a.logS;
  constexpr int NT = tile_threads(K, LT);
  constexpr u32 nslot = 1u << K;
  constexpr int NR = (K + 3) / 4;
  uint2* tw = reinterpret_cast<uint2*>(smem);
  u32* data = smem + tw_words(K);
  const u32 tile_elems = nslot << logT;
  const u32* __restrict__ in = a.in + (size_t)blockIdx.y * a.in_stride + (size_t)blockIdx.z * a.in_coset_stride;
  u32* __restrict__ out = a.out + (size_t)blockIdx.y * a.out_stride + (size_t)blockIdx.z * a.out_coset_stride;
  const uint2* __restrict__ four = a.four;
  fill_small_twiddles<K>(tw, a.inverse, a.small_tw);
  const u32 t0 = blockIdx.x << logT;   // offsets t0 .. t0+T-1

  // both kinds run the DIF network: natural slots in, bit-reversed slots out.  For the final DIT
  // level the natural slot s is fed from row bitrev(s) (pre-twiddled), see the file header.
  auto gload = [&](u32 slot, u32 q) {
    const u32 row = FINAL ? bitrev32(slot, K) : slot;
    const size_t pos = ((size_t)row << logS) + t0 + q;
    const u32 v = in[pos];
    return FINAL ? shoup_mul(v, __ldg(four + pos)) : fp_raw(v);
  };
  auto gstore = [&](u32 slot, u32 q, Fp v) {      // DIF only: post-twiddle, in-place position
    const size_t pos = ((size_t)slot << logS) + t0 + q;
    out[pos] = shoup_mul(v.v, __ldg(four + pos)).v;
  };
  auto sload = [&](u32 slot, u32 q) { return fp_raw(data[saddr(slot, q, logT)]); };
  auto sstore = [&](u32 slot, u32 q, Fp v) { data[saddr(slot, q, logT)] = v.v; };

  __syncthreads();   // twiddles visible
  constexpr int R0 = round_bits(K, 0);
  if constexpr (NR == 1) {
    if constexpr (FINAL) round_io<K, R0, 0, true, false, NT>(logT, tw, gload, sstore);
    else round_io<K, R0, 0, true, false, NT>(logT, tw, gload, gstore);
  } else {
    round_io<K, R0, round_lo(K, 0), true, false, NT>(logT, tw, gload, sstore);
    __syncthreads();
    constexpr int R1 = round_bits(K, 1);
    if constexpr (NR == 2) {
      if constexpr (FINAL) round_io<K, R1, 0, true, false, NT>(logT, tw, sload, sstore);
      else round_io<K, R1, 0, true, false, NT>(logT, tw, sload, gstore);
    } else {
      round_io<K, R1, round_lo(K, 1), true, false, NT>(logT, tw, sload, sstore);
      __syncthreads();
      constexpr int R2 = round_bits(K, 2);
      if constexpr (FINAL) round_io<K, R2, 0, true, false, NT>(logT, tw, sload, sstore);
      else round_io<K, R2, 0, true, false, NT>(logT, tw, sload, gstore);
    }
  }
  if constexpr (FINAL) {
    __syncthreads();
    // slot o holds k_hi = bitrev_K(o); natural index k_hi*S + t lands at bitrev_logS(t) * 2^K + o:
    // transposed through shared memory so that the stores run along o
    for (u32 i = threadIdx.x; i < tile_elems; i += blockDim.x) {
      const u32 o = i & (nslot - 1), q = i >> K;
      const size_t pos = ((size_t)bitrev32(t0 + q, logS) << K) + o;
      out[pos] = data[saddr(o, q, logT)];
    }
  }
}

template <int K, int LT>
__global__ void __launch_bounds__(1024, 1) ntt_strided_kernel(StridedArgs a) {
  extern __shared__ __align__(16) u32 smem[];
  if (a.final_dit) strided_body<K, true, LT>(a, smem);
  else strided_body<K, false, LT>(a, smem);
}
// default tile width of the strided level: 2^13 elements per tile, at most 32 offsets per slot
static constexpr int STRIDED_TILE_LOG = 13;
__host__ __device__ constexpr int strided_default_lt(int K) { return STRIDED_TILE_LOG - K > 5 ? 5 : (STRIDED_TILE_LOG - K < 0 ? 0 : STRIDED_TILE_LOG - K); }

}  // namespace zkb
#include "ntt_lean.cuh"
namespace zkb {

// ---- contiguous level (optionally fused inverse -> scale -> forward per coset) --------------------
struct ContigArgs {
  const u32* in; u32* out;
  size_t in_stride, out_stride;      // elements between columns
  const uint2* small_tw;
  int K, logT, logn;                 // groups of 2^K contiguous elements; column length 2^logn
  u32 total_groups;                  // ncols * 2^(logn-K)
  int mode;                          // 0: DIF only (inverse flag applies)
                                     // 1: DIF(inverse) -> scale_c -> DIT(forward), per coset
  int inverse;
  int ncoset;
  int bitrev_store;                  // mode 1, single level (K == logn): store bit-reversed rows
  const uint2* scale;                // per coset: n Shoup pairs, shift_c^bitrev(p) / n at position p
  size_t out_coset_stride;           // element offset between coset outputs
};

static constexpr int CONTIG_TILE_LOG = 12;
__host__ __device__ constexpr int contig_default_lt(int K) { return CONTIG_TILE_LOG - K < 0 ? 0 : CONTIG_TILE_LOG - K; }

template <int K, int LT>
__global__ void __launch_bounds__(512, 2) ntt_contig_kernel(ContigArgs a) {
  extern __shared__ __align__(16) u32 smem[];
  const int logT = LT >= 0 ? LT : a.logT;
  constexpr int NT = tile_threads(K, LT);
  constexpr u32 nslot = 1u << K;
  constexpr int NR = (K + 3) / 4;
  const u32 T = 1u << logT;
  constexpr u32 ldg = nslot + (nslot >> 3) + 1;
  uint2* tw_a = reinterpret_cast<uint2*>(smem);                  // first transform's twiddles
  uint2* tw_b = reinterpret_cast<uint2*>(smem + tw_words(K));    // forward twiddles for the fused DIT
  u32* bufA = smem + 2 * tw_words(K);
  u32* bufB = bufA + (size_t)T * ldg;
  const u32 tile_elems = nslot << logT;
  const int sub_bits = a.logn - K;                 // groups per column = 2^sub_bits
  fill_small_twiddles<K>(tw_a, a.mode == 1 ? true : (a.inverse != 0), a.small_tw);
  if (a.mode == 1) fill_small_twiddles<K>(tw_b, false, a.small_tw);
  const u32 f0 = blockIdx.x << logT;
  const u32* __restrict__ gin = a.in;

  auto gaddr = [&](u32 slot, u32 q, size_t stride) {     // element offset of (group f0+q, slot)
    const u32 f = f0 + q;
    return (size_t)(f >> sub_bits) * stride + ((size_t)(f & ((1u << sub_bits) - 1)) << K) + slot;
  };
  auto gload = [&](u32 slot, u32 q) { return fp_raw((f0 + q) < a.total_groups ? gin[gaddr(slot, q, a.in_stride)] : 0u); };
  auto aload = [&](u32 slot, u32 q) { return fp_raw(bufA[caddr(slot, q, ldg)]); };
  auto astore = [&](u32 slot, u32 q, Fp v) { bufA[caddr(slot, q, ldg)] = v.v; };
  auto bload = [&](u32 slot, u32 q) { return fp_raw(bufB[caddr(slot, q, ldg)]); };
  auto bstore = [&](u32 slot, u32 q, Fp v) { bufB[caddr(slot, q, ldg)] = v.v; };

  __syncthreads();   // twiddles visible
  // ---- DIF (top bits first); the last round's result is left in shared memory (bufA) ----
  constexpr int R0 = round_bits(K, 0);
  round_io<K, R0, round_lo(K, 0), true, true, NT>(logT, tw_a, gload, astore);
  __syncthreads();
  if constexpr (NR > 1) {
    constexpr int R1 = round_bits(K, 1);
    round_io<K, R1, round_lo(K, 1), true, true, NT>(logT, tw_a, aload, astore);
    __syncthreads();
    if constexpr (NR > 2) {
      constexpr int R2 = round_bits(K, 2);
      round_io<K, R2, round_lo(K, 2), true, true, NT>(logT, tw_a, aload, astore);
      __syncthreads();
    }
  }

  if (a.mode == 0) {
    for (u32 i = threadIdx.x; i < tile_elems; i += blockDim.x) {
      const u32 slot = i & (nslot - 1), q = i >> K;
      if (f0 + q >= a.total_groups) continue;
      a.out[gaddr(slot, q, a.out_stride)] = bufA[caddr(slot, q, ldg)];
    }
    return;
  }

  for (int c = 0; c < a.ncoset; c++) {
    const uint2* __restrict__ sc = a.scale + ((size_t)c << a.logn);
    u32* __restrict__ outc = a.out + (size_t)c * a.out_coset_stride;
    // coefficient at bit-reversed position p gets shift_c^bitrev(p) / n, folded into the load of
    // the first DIT round (bottom bits), which reads the coefficients from bufA
    auto cload = [&](u32 slot, u32 q) {
      const u32 sub = (f0 + q) & ((1u << sub_bits) - 1);
      return shoup_mul(bufA[caddr(slot, q, ldg)], __ldg(sc + (((size_t)sub << K) + slot)));
    };
    auto gstore = [&](u32 slot, u32 q, Fp v) {
      if (f0 + q >= a.total_groups) return;
      outc[gaddr(slot, q, a.out_stride)] = v.v;
    };
    // ---- DIT: the same bit ranges bottom-up; the last round stores straight to global memory ----
    constexpr int RL = round_bits(K, NR - 1);
    if constexpr (NR == 1) {
      if (a.bitrev_store) round_io<K, RL, 0, false, true, NT>(logT, tw_b, cload, bstore);
      else round_io<K, RL, 0, false, true, NT>(logT, tw_b, cload, gstore);
    } else {
      round_io<K, RL, 0, false, true, NT>(logT, tw_b, cload, bstore);
      __syncthreads();
      if constexpr (NR == 2) {
        if (a.bitrev_store) round_io<K, R0, round_lo(K, 0), false, true, NT>(logT, tw_b, bload, bstore);
        else round_io<K, R0, round_lo(K, 0), false, true, NT>(logT, tw_b, bload, gstore);
      } else {
        constexpr int R1 = round_bits(K, 1);
        round_io<K, R1, round_lo(K, 1), false, true, NT>(logT, tw_b, bload, bstore);
        __syncthreads();
        if (a.bitrev_store) round_io<K, R0, round_lo(K, 0), false, true, NT>(logT, tw_b, bload, bstore);
        else round_io<K, R0, round_lo(K, 0), false, true, NT>(logT, tw_b, bload, gstore);
      }
    }
    if (a.bitrev_store) {
      // single level (the whole column is one group): rows leave in bit-reversed order
      __syncthreads();
      for (u32 i = threadIdx.x; i < tile_elems; i += blockDim.x) {
        const u32 o = i & (nslot - 1), q = i >> K;
        if (f0 + q >= a.total_groups) continue;
        outc[gaddr(o, q, a.out_stride)] = bufB[caddr(bitrev32(o, K), q, ldg)];
      }
    }
    __syncthreads();
  }
}

// ---- launch helpers ------------------------------------------------------------------------------

// Dynamic shared memory above 48 KB is a per-DEVICE function attribute: set for every instantiation
// when a context initialises its tables on its device (NttTables::init), not behind process-wide flags.
template <int KK>
static void set_ntt_attrs_from() {
  ZKB_CUDA(cudaFuncSetAttribute(ntt_strided_kernel<KK, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ZKB_CUDA(cudaFuncSetAttribute(ntt_strided_kernel<KK, strided_default_lt(KK)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  ZKB_CUDA(cudaFuncSetAttribute(ntt_contig_kernel<KK, -1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  if constexpr (KK < KMAX) set_ntt_attrs_from<KK + 1>();
}
void ntt_set_lean_attributes();
void ntt_set_device_attributes() { set_ntt_attrs_from<1>(); ntt_set_lean_attributes(); }

static void split_levels(int logn, int& K1, int& K2) {
  if (logn <= KMAX) { K1 = 0; K2 = logn; return; }
  // size of the contiguous level, fastest measured split per log n (tools/k2_sweep.py on a B200 with the
  // lean strided kernels, profiles/r02_k2_sweep.jsonl): the contiguous level likes 2^10 / 2^11 (three
  // radix rounds of a 4096-element tile), the strided level two radix-16 rounds
  static const int best_k2[25] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 9, 9, 9, 9, 10, 11, 11, 11, 11, 11, 12, 12};
  K2 = best_k2[logn];
  // experiment knob: size of the contiguous level (e.g. 12: 2^18 = 2^6 x 2^12, five radix rounds instead of six)
  static const int env_k2 = env_int("ZKB200_NTT_K2", 0);
  const int force_k2 = g_ntt_force_k2.load() ? g_ntt_force_k2.load() : env_k2;
  if (force_k2 > 0 && force_k2 <= KMAX && logn - force_k2 >= 1 && logn - force_k2 <= KMAX) K2 = force_k2;
  K1 = logn - K2;
  if (K2 > KMAX) throw std::runtime_error("zkb200: NTT size out of range");
}

std::atomic<int> g_ntt_lean{3};      // zkb200_set_option("ntt_lean"): bit 0 lean strided kernels, bit 1 lean contiguous kernel

template <int K, int LOGS>
static void launch_lean_pair(const LeanStridedArgs& a, bool final_dit, dim3 grid, cudaStream_t s) {
  if (final_dit) ntt_strided_lean_kernel<K, LOGS, true><<<grid, 1 << (K + 1), lean_strided_smem<K>(true), s>>>(a);
  else ntt_strided_lean_kernel<K, LOGS, false><<<grid, 1 << (K + 1), lean_strided_smem<K>(false), s>>>(a);
}
template <int K, int LOGS>
static void set_lean_attrs_pair() {
  ZKB_CUDA(cudaFuncSetAttribute(ntt_strided_lean_kernel<K, LOGS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lean_strided_smem<K>(true)));
  ZKB_CUDA(cudaFuncSetAttribute(ntt_strided_lean_kernel<K, LOGS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lean_strided_smem<K>(false)));
}
// the (level size, stride) pairs the two-level splits of 2^13 .. 2^21 use
#define ZKB_LEAN_PAIRS(X) \
  X(5, 8) X(6, 8) X(7, 8) X(8, 8) X(4, 9) X(5, 9) X(6, 9) X(7, 9) X(8, 9) X(9, 9) \
  X(4, 10) X(5, 10) X(6, 10) X(7, 10) X(8, 10) X(9, 10) X(4, 11) X(5, 11) X(6, 11) X(7, 11) X(8, 11) X(9, 11) \
  X(6, 12) X(7, 12) X(8, 12) X(9, 12)
static bool launch_strided_lean(const LeanStridedArgs& a, int K, int logS, bool final_dit, dim3 grid, cudaStream_t s) {
#define X(KK, SS) if (K == KK && logS == SS) { launch_lean_pair<KK, SS>(a, final_dit, grid, s); return true; }
  ZKB_LEAN_PAIRS(X)
#undef X
  return false;
}
template <int K>
static void set_lean_contig_attr() {
  ZKB_CUDA(cudaFuncSetAttribute(ntt_contig_lean_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lean_contig_smem<K>()));
}
void ntt_set_lean_attributes() {
#define X(KK, SS) set_lean_attrs_pair<KK, SS>();
  ZKB_LEAN_PAIRS(X)
#undef X
  set_lean_contig_attr<8>(); set_lean_contig_attr<9>(); set_lean_contig_attr<10>(); set_lean_contig_attr<11>(); set_lean_contig_attr<12>();
}
// the fused middle launch of a two-level coset LDE with the lean kernel; false: not covered (generic kernel)
static bool launch_contig_lean(const NttTables& tb, const u32* in, size_t in_stride, u32* out, size_t out_stride, size_t out_coset_stride,
                               const uint2* scale, int ncoset, int K, int logn, size_t ncols, cudaStream_t s) {
  if (!(g_ntt_lean.load() & 2) || K < 8 || K > 12 || logn < 13) return false;
  LeanContigArgs a;
  a.in = in; a.out = out; a.in_stride = in_stride; a.out_stride = out_stride; a.out_coset_stride = out_coset_stride;
  a.small_tw = (const uint2*)tb.small_tw; a.scale = scale; a.ncoset = ncoset;
  dim3 grid(1u << (logn - 12), (unsigned)ncols);
  switch (K) {
    case 8: ntt_contig_lean_kernel<8><<<grid, 256, lean_contig_smem<8>(), s>>>(a, logn); break;
    case 9: ntt_contig_lean_kernel<9><<<grid, 256, lean_contig_smem<9>(), s>>>(a, logn); break;
    case 10: ntt_contig_lean_kernel<10><<<grid, 256, lean_contig_smem<10>(), s>>>(a, logn); break;
    case 11: ntt_contig_lean_kernel<11><<<grid, 256, lean_contig_smem<11>(), s>>>(a, logn); break;
    default: ntt_contig_lean_kernel<12><<<grid, 256, lean_contig_smem<12>(), s>>>(a, logn); break;
  }
  ZKB_CHECK_LAUNCH();
  return true;
}

static void launch_strided(const NttTables& tb, const u32* in, size_t in_stride, u32* out, size_t out_stride, size_t ncols,
                           int K, int logS, bool inverse, bool final_dit, int ncoset, size_t in_coset_stride,
                           size_t out_coset_stride, cudaStream_t s) {
  if (g_ntt_lean.load() & 1) {
    LeanStridedArgs la;
    la.in = in; la.out = out; la.in_stride = in_stride; la.out_stride = out_stride;
    la.in_coset_stride = in_coset_stride; la.out_coset_stride = out_coset_stride;
    la.small_tw = (const uint2*)tb.small_tw;
    la.inverse = inverse ? 1 : 0;
    if (logS >= 5) {
      dim3 grid(1u << (logS - 5), (unsigned)ncols, (unsigned)ncoset);
      // the table lookup happens only for pairs that exist: probe with a null launch config first
      bool have = false;
#define X(KK, SS) if (K == KK && logS == SS) have = true;
      ZKB_LEAN_PAIRS(X)
#undef X
      if (have) {
        la.four = (const uint2*)tb.four_step_table(K, logS, inverse, s);
        launch_strided_lean(la, K, logS, final_dit, grid, s);
        ZKB_CHECK_LAUNCH();
        return;
      }
    }
  }
  StridedArgs a;
  a.in = in; a.out = out; a.in_stride = in_stride; a.out_stride = out_stride;
  a.small_tw = (const uint2*)tb.small_tw;
  a.four = (const uint2*)tb.four_step_table(K, logS, inverse, s);
  a.logS = logS; a.inverse = inverse ? 1 : 0; a.final_dit = final_dit ? 1 : 0;
  a.in_coset_stride = in_coset_stride; a.out_coset_stride = out_coset_stride;
  static const int tile_log = env_int("ZKB200_NTT_STRIDED_TILE_LOG", STRIDED_TILE_LOG);
  int logT = tile_log - K;                       // 8192 elements (512 threads) per tile: two CTAs per SM
  if (logT > 5) logT = 5;
  if (logT > logS) logT = logS;
  if (logT < 0) logT = 0;
  a.logT = logT;
  const size_t tile_elems = (size_t)1 << (K + logT);
  unsigned threads = (unsigned)std::max<size_t>(32, tile_elems / ELEMS_PER_THREAD);
  size_t smem = ((size_t)tw_words(K) + ((size_t)1 << (K + logT))) * sizeof(u32);
  dim3 grid(1u << (logS - logT), (unsigned)ncols, (unsigned)ncoset);
#define ZKB_STRIDED_CASE(KK)                                                                                              \
  case KK:                                                                                                                \
    if (logT == strided_default_lt(KK)) ntt_strided_kernel<KK, strided_default_lt(KK)><<<grid, threads, smem, s>>>(a);    \
    else ntt_strided_kernel<KK, -1><<<grid, threads, smem, s>>>(a);                                                       \
    break;
  switch (K) {
    ZKB_STRIDED_CASE(1) ZKB_STRIDED_CASE(2) ZKB_STRIDED_CASE(3) ZKB_STRIDED_CASE(4) ZKB_STRIDED_CASE(5) ZKB_STRIDED_CASE(6)
    ZKB_STRIDED_CASE(7) ZKB_STRIDED_CASE(8) ZKB_STRIDED_CASE(9) ZKB_STRIDED_CASE(10) ZKB_STRIDED_CASE(11) ZKB_STRIDED_CASE(12)
    default: throw std::runtime_error("zkb200: strided NTT level out of range");
  }
#undef ZKB_STRIDED_CASE
  ZKB_CHECK_LAUNCH();
}

static void launch_contig(const NttTables& tb, ContigArgs a, size_t ncols, cudaStream_t s) {
  a.small_tw = (const uint2*)tb.small_tw;
  const int K = a.K;
  const size_t groups = ncols << (a.logn - K);
  a.total_groups = (u32)groups;
  static const int ctile_log = env_int("ZKB200_NTT_CONTIG_TILE_LOG", CONTIG_TILE_LOG);
  int logT = ctile_log - K;
  if (logT < 0) logT = 0;
  while (logT > 0 && ((size_t)1 << logT) > groups) logT--;
  a.logT = logT;
  const size_t tile_elems = (size_t)1 << (K + logT);
  unsigned threads = (unsigned)std::max<size_t>(32, tile_elems / ELEMS_PER_THREAD);
  const size_t ldg = ((size_t)1 << K) + (((size_t)1 << K) >> 3) + 1;
  size_t smem = ((size_t)2 * tw_words(K) + ((size_t)(a.mode == 1 ? 2 : 1) << logT) * ldg) * sizeof(u32);
  unsigned grid = (unsigned)((groups + ((size_t)1 << logT) - 1) >> logT);
#define ZKB_CONTIG_CASE(KK)                                                                                               \
  case KK:                                                                                                                \
    ntt_contig_kernel<KK, -1><<<grid, threads, smem, s>>>(a);   /* fixed tile width: no measurable gain here */          \
    break;
  switch (K) {
    ZKB_CONTIG_CASE(1) ZKB_CONTIG_CASE(2) ZKB_CONTIG_CASE(3) ZKB_CONTIG_CASE(4) ZKB_CONTIG_CASE(5) ZKB_CONTIG_CASE(6)
    ZKB_CONTIG_CASE(7) ZKB_CONTIG_CASE(8) ZKB_CONTIG_CASE(9) ZKB_CONTIG_CASE(10) ZKB_CONTIG_CASE(11) ZKB_CONTIG_CASE(12)
    default: throw std::runtime_error("zkb200: contiguous NTT level out of range");
  }
#undef ZKB_CONTIG_CASE
  ZKB_CHECK_LAUNCH();
}

// scale[c][p] = shift_c^bitrev(p) / n as Shoup pairs; output block c holds coset shift * w_N^bitrev(c)
__global__ void scale_table_kernel(uint2* out, int logn, int log_blowup, Fp shift, Fp ninv, Fp wN) {
  const size_t n = (size_t)1 << logn;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= (n << log_blowup)) return;
  u32 c = (u32)(i >> logn), p = (u32)i & (u32)(n - 1);
  Fp sh = shift * fp_pow(wN, bitrev32(c, log_blowup));
  out[i] = shoup_pair(ninv * fp_pow(sh, bitrev32(p, logn)));
}
std::shared_ptr<void> NttTables::scale_table(unsigned log_n, unsigned log_blowup, Fp shift, cudaStream_t s) const {
  std::lock_guard<std::mutex> lock(mu);
  const u64 key = ((u64)shift.v << 16) | (log_n << 4) | log_blowup;
  auto it = scale_cache.find(key);
  if (it != scale_cache.end()) return it->second;
  const size_t count = (size_t)1 << (log_n + log_blowup);
  if (scale_cache_bytes + count * sizeof(uint2) > ((size_t)1 << 30)) {
    // drop the cache's references only: a table some lane still uses lives on in that lane's
    // keep-alive list and is freed when the lane lets go of it
    scale_cache.clear(); scale_bytes.clear(); scale_cache_bytes = 0;
  }
  void* p = nullptr;
  ZKB_CUDA(cudaMalloc(&p, count * sizeof(uint2)));
  std::shared_ptr<void> sp(p, [](void* q) { cudaFree(q); });
  Fp ninv = fp_inv(fp_from_canonical((u32)(((size_t)1 << log_n) % KB_P)));
  scale_table_kernel<<<ceil_div(count, 256), 256, 0, s>>>((uint2*)p, (int)log_n, (int)log_blowup, shift, ninv,
                                                         two_adic_generator(log_n + log_blowup));
  ZKB_CHECK_LAUNCH();
  ZKB_CUDA(cudaStreamSynchronize(s));   // published to every lane: must be complete
  scale_cache[key] = sp;
  scale_bytes[key] = count * sizeof(uint2);
  scale_cache_bytes += count * sizeof(uint2);
  return sp;
}

void coset_lde_batch(const NttTables& tb, const u32* in, size_t in_stride, u32* out, size_t out_stride,
                     unsigned log_n, size_t width, unsigned log_blowup, Fp shift, cudaStream_t s, KeepAlive& keep) {
  if (width == 0) return;
  const size_t n = (size_t)1 << log_n;
  const unsigned ncoset = 1u << log_blowup;
  if (log_n == 0) {
    for (unsigned c = 0; c < ncoset; c++)
      ZKB_CUDA(cudaMemcpy2DAsync(out + c, out_stride * 4, in, in_stride * 4, 4, width, cudaMemcpyDeviceToDevice, s));
    return;
  }
  int K1, K2;
  split_levels((int)log_n, K1, K2);
  keep.push_back(tb.scale_table(log_n, log_blowup, shift, s));
  const uint2* scale = (const uint2*)keep.back().get();
  // column chunks bound the scratch (n x chunk x (1 + cosets) words).  The kernels are
  // instruction-issue bound, not bandwidth bound, so large chunks (fewer, fuller launches) beat
  // L2-resident ones: measured 3.3 ms at 2^26 vs 4.3 ms at 2^22 for 2^18 x 512 (profiles/README.md)
  const size_t chunk_elems = (size_t)1 << std::min(30, std::max(10, env_int("ZKB200_NTT_CHUNK_LOG", 26)));
  size_t chunk = chunk_elems >> log_n;
  if (chunk < 1) chunk = 1;
  if (chunk > width) chunk = width;
  if (chunk > 32768) chunk = 32768;
  const bool two_level = K1 > 0;
  DevBuf half(two_level ? chunk * n : 0, s), xbuf(two_level ? chunk * n * ncoset : 0, s);
  for (size_t c0 = 0; c0 < width; c0 += chunk) {
    const size_t nc = width - c0 < chunk ? width - c0 : chunk;
    ContigArgs b;
    b.K = K2; b.logn = (int)log_n; b.mode = 1; b.inverse = 1; b.ncoset = (int)ncoset;
    b.scale = scale;
    if (!two_level) {
      b.in = in + c0 * in_stride; b.in_stride = in_stride;
      b.out = out + c0 * out_stride; b.out_stride = out_stride; b.out_coset_stride = n;
      b.bitrev_store = 1;
      launch_contig(tb, b, nc, s);
      continue;
    }
    launch_strided(tb, in + c0 * in_stride, in_stride, half.p, n, nc, K1, K2, true, false, 1, 0, 0, s);
    b.in = half.p; b.in_stride = n;
    b.out = xbuf.p; b.out_stride = n; b.out_coset_stride = chunk * n;
    b.bitrev_store = 0;
    if (!launch_contig_lean(tb, half.p, n, xbuf.p, n, chunk * n, scale, (int)ncoset, K2, (int)log_n, nc, s)) launch_contig(tb, b, nc, s);
    launch_strided(tb, xbuf.p, n, out + c0 * out_stride, out_stride, nc, K1, K2, false, true, (int)ncoset, chunk * n, n, s);
  }
}

__global__ void bitrev_rows_kernel(const u32* __restrict__ in, u32* __restrict__ out, unsigned log_n) {
  size_t n = (size_t)1 << log_n;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u32* ci = in + blockIdx.y * n;
  u32* co = out + blockIdx.y * n;
  co[i] = ci[bitrev32((u32)i, log_n)];
}
void bitrev_rows(const u32* in, u32* out, unsigned log_n, size_t width, cudaStream_t s) {
  if (!width) return;
  size_t n = (size_t)1 << log_n;
  for (size_t c0 = 0; c0 < width; c0 += 32768) {
    size_t nc = width - c0 < 32768 ? width - c0 : 32768;
    dim3 grid(ceil_div(n, 256), (unsigned)nc);
    bitrev_rows_kernel<<<grid, 256, 0, s>>>(in + c0 * n, out + c0 * n, log_n);
    ZKB_CHECK_LAUNCH();
  }
}

__global__ void scale_all_kernel(u32* d, size_t count, Fp c) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < count) d[i] = (fp_raw(d[i]) * c).v;
}

void ntt_batch(const NttTables& tb, const u32* in, u32* out, unsigned log_n, size_t width, bool inverse, bool bitrev_out,
               cudaStream_t s) {
  if (!width) return;
  const size_t n = (size_t)1 << log_n;
  if (log_n == 0) { ZKB_CUDA(cudaMemcpyAsync(out, in, width * 4, cudaMemcpyDeviceToDevice, s)); return; }
  int K1, K2;
  split_levels((int)log_n, K1, K2);
  for (size_t c0 = 0; c0 < width; c0 += 32768) {
    size_t nc = width - c0 < 32768 ? width - c0 : 32768;
    DevBuf tmp(bitrev_out ? 0 : nc * n, s);
    u32* dst = bitrev_out ? out + c0 * n : tmp.p;
    const u32* src = in + c0 * n;
    if (K1 > 0) {
      launch_strided(tb, src, n, dst, n, nc, K1, K2, inverse, false, 1, 0, 0, s);
      src = dst;
    }
    ContigArgs b;
    b.in = src; b.in_stride = n; b.out = dst; b.out_stride = n; b.out_coset_stride = 0;
    b.K = K2; b.logn = (int)log_n; b.mode = 0; b.inverse = inverse ? 1 : 0; b.ncoset = 1; b.bitrev_store = 0;
    b.scale = nullptr;
    launch_contig(tb, b, nc, s);
    if (!bitrev_out) bitrev_rows(tmp.p, out + c0 * n, log_n, nc, s);
  }
  if (inverse) {
    Fp ninv = fp_inv(fp_from_canonical((u32)(n % KB_P)));
    size_t count = n * width;
    scale_all_kernel<<<ceil_div(count, 256), 256, 0, s>>>(out, count, ninv);
    ZKB_CHECK_LAUNCH();
  }
}

}  // namespace zkb

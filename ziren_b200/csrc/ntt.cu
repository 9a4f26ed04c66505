// K1: batched KoalaBear NTT / coset LDE, column-major.
//
// A transform of length N = 2^logN is a sequence of PASSES; each pass runs small NTTs of size
// 2^K inside shared-memory tiles (2^K slots x T neighbouring offsets of one column) and applies
// the four-step twiddle w_L^(i*t) that links it to the next level.  Three pass kinds:
//   DIF        natural -> bit-reversed, in place         (inverse transform of the evaluations)
//   DIT        bit-reversed -> natural, in place         (forward transform, inner levels)
//   DIT_FINAL  last forward level, stores BIT-REVERSED   (the committed LDE row order)
// so the LDE never needs a separate permutation or transpose pass: evaluations (natural)
// --DIF,w^-1--> coefficients (bit-reversed) --scale by shift^i / n on load, DIT--> coset
// evaluations written straight into their bit-reversed rows.  Columns are processed in chunks
// sized to stay L2-resident (126 MB on B200) between the passes.
#include "ntt.h"
#include <algorithm>

namespace zkb {

static constexpr int TW_SPLIT = 12;
static constexpr int MAX_TILE_LOG = 13;   // 8192 elements = 32 KB (+ padding) per CTA
static constexpr int NTT_THREADS = 256;

__global__ void tw_init_kernel(u32* lo, u32* hi) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (1 << TW_SPLIT)) return;
  Fp w = two_adic_generator(24);
  lo[i] = fp_pow(w, (u64)i).v;
  hi[i] = fp_pow(w, (u64)i << TW_SPLIT).v;
}
void NttTables::init(cudaStream_t s) {
  ZKB_CUDA(cudaMalloc((void**)&tw_lo, sizeof(u32) << TW_SPLIT));
  ZKB_CUDA(cudaMalloc((void**)&tw_hi, sizeof(u32) << TW_SPLIT));
  tw_init_kernel<<<(1 << TW_SPLIT) / 256, 256, 0, s>>>(tw_lo, tw_hi);
  ZKB_CHECK_LAUNCH();
}
void NttTables::destroy() {
  if (tw_lo) cudaFree(tw_lo);
  if (tw_hi) cudaFree(tw_hi);
  tw_lo = tw_hi = nullptr;
}

// w^E for E in [0, 2^24)
__device__ __forceinline__ Fp tw_pow(const u32* __restrict__ lo, const u32* __restrict__ hi, u32 E) {
  return fp_raw(__ldg(hi + (E >> TW_SPLIT))) * fp_raw(__ldg(lo + (E & ((1u << TW_SPLIT) - 1))));
}

enum PassKind { PASS_DIF = 0, PASS_DIT = 1, PASS_DIT_FINAL = 2 };

struct PassArgs {
  const u32* in;
  u32* out;
  size_t in_stride, out_stride;  // elements between columns
  const u32* tw_lo;
  const u32* tw_hi;
  const u32* scale_a;            // optional load scaling: f(p) = a[p & mask] * b[p >> split]
  const u32* scale_b;
  int scale_split;
  int K, logS, logT, logN;
  int kind, inverse, contig;
};

__global__ void __launch_bounds__(NTT_THREADS) ntt_pass_kernel(PassArgs a) {
  extern __shared__ u32 smem[];
  const int K = a.K, logT = a.logT, logS = a.logS;
  const u32 nslot = 1u << K, T = 1u << logT;
  const u32 half_n = nslot >> 1;
  u32* tw = smem;                      // 2^(K-1) small twiddles
  u32* data = smem + (half_n ? half_n : 1);
  const u32 row = a.contig ? (nslot + 1) : (T + 1);   // padded leading dimension
  const u32 tile_elems = nslot << logT;
  const u32* __restrict__ in = a.in + (size_t)blockIdx.y * a.in_stride;
  u32* __restrict__ out = a.out + (size_t)blockIdx.y * a.out_stride;
  const u32 logL = K + logS;
  const u32 mask24 = (1u << 24) - 1;

  // small twiddles w_{2^K}^(+-m)
  for (u32 m = threadIdx.x; m < half_n; m += blockDim.x) {
    u32 E = m << (24 - K);
    if (a.inverse) E = (0u - E) & mask24;
    tw[m] = tw_pow(a.tw_lo, a.tw_hi, E).v;
  }

  // tile origin: flat offset index f0 = tile * T enumerates (sub-problem, t) pairs
  const u32 f0 = blockIdx.x << logT;
  u32 base, t0;
  if (a.contig) { base = f0 << K; t0 = 0; }            // S == 1: T consecutive groups
  else { u32 sub = f0 >> logS; t0 = f0 & ((1u << logS) - 1); base = (sub << logL) + t0; }

  // ---- load (+ optional scaling, + pre-twiddle for DIT kinds) ----
  for (u32 i = threadIdx.x; i < tile_elems; i += blockDim.x) {
    u32 j, q, pos;
    if (a.contig) { j = i & (nslot - 1); q = i >> K; pos = base + i; }
    else { q = i & (T - 1); j = i >> logT; pos = base + (j << logS) + q; }
    Fp v = fp_raw(in[pos]);
    if (a.scale_a) {
      Fp f = fp_raw(__ldg(a.scale_a + (pos & ((1u << a.scale_split) - 1)))) * fp_raw(__ldg(a.scale_b + (pos >> a.scale_split)));
      v = v * f;
    }
    u32 slot = j;
    if (a.kind != PASS_DIF) {
      u32 ilo = bitrev32(j, K);
      u32 t = a.contig ? 0 : (t0 + q);
      if (logS) {
        u32 e = (ilo * t) & ((1u << logL) - 1);
        u32 E = e << (24 - logL);
        if (a.inverse) E = (0u - E) & mask24;
        v = v * tw_pow(a.tw_lo, a.tw_hi, E);
      }
      if (a.kind == PASS_DIT_FINAL) slot = ilo;
    }
    data[a.contig ? (q * row + slot) : (slot * row + q)] = v.v;
  }
  __syncthreads();

  // ---- butterfly network over the slot dimension ----
  const u32 n_bf = half_n << logT;
  const bool dif = (a.kind != PASS_DIT);
  for (int s = 0; s < K; s++) {
    const u32 lh = dif ? (K - 1 - s) : s;       // log2(half)
    const u32 half = 1u << lh;
    for (u32 i = threadIdx.x; i < n_bf; i += blockDim.x) {
      u32 bf, q;
      if (a.contig) { bf = i & (half_n - 1); q = i >> (K - 1); }
      else { q = i & (T - 1); bf = i >> logT; }
      u32 k = bf & (half - 1);
      u32 j0 = ((bf >> lh) << (lh + 1)) + k, j1 = j0 + half;
      u32 i0 = a.contig ? (q * row + j0) : (j0 * row + q);
      u32 i1 = a.contig ? (q * row + j1) : (j1 * row + q);
      Fp w = fp_raw(tw[k << (K - 1 - lh)]);
      Fp x = fp_raw(data[i0]), y = fp_raw(data[i1]);
      if (dif) { data[i0] = (x + y).v; data[i1] = ((x - y) * w).v; }
      else { Fp yw = y * w; data[i0] = (x + yw).v; data[i1] = (x - yw).v; }
    }
    __syncthreads();
  }

  // ---- store (+ post-twiddle for DIF) ----
  if (a.kind == PASS_DIT_FINAL) {
    // slot o holds output k_hi = bitrev_K(o); natural index k_hi * S + t lands at
    // bitrev_logN = bitrev_logS(t) * 2^K + o  -> contiguous runs over o
    for (u32 i = threadIdx.x; i < tile_elems; i += blockDim.x) {
      u32 o = i & (nslot - 1), q = i >> K;
      u32 t = a.contig ? 0 : (t0 + q);
      u32 pos = (bitrev32(t, logS) << K) + o;
      if (a.contig) pos += (f0 + q) << K;  // only reachable with a single group (logS == 0, logN == K)
      out[pos] = data[a.contig ? (q * row + o) : (o * row + q)];
    }
  } else {
    for (u32 i = threadIdx.x; i < tile_elems; i += blockDim.x) {
      u32 o, q, pos;
      if (a.contig) { o = i & (nslot - 1); q = i >> K; pos = base + i; }
      else { q = i & (T - 1); o = i >> logT; pos = base + (o << logS) + q; }
      Fp v = fp_raw(data[a.contig ? (q * row + o) : (o * row + q)]);
      if (a.kind == PASS_DIF && logS) {
        u32 k1 = bitrev32(o, K);
        u32 e = (k1 * (t0 + q)) & ((1u << logL) - 1);
        u32 E = e << (24 - logL);
        if (a.inverse) E = (0u - E) & mask24;
        v = v * tw_pow(a.tw_lo, a.tw_hi, E);
      }
      out[pos] = v.v;
    }
  }
}

static void launch_pass(const NttTables& tb, const u32* in, size_t in_stride, u32* out, size_t out_stride,
                        size_t ncols, int logN, int K, int logS, int kind, bool inverse, const u32* scale_a,
                        const u32* scale_b, int scale_split, cudaStream_t s) {
  PassArgs a;
  a.in = in; a.out = out; a.in_stride = in_stride; a.out_stride = out_stride;
  a.tw_lo = tb.tw_lo; a.tw_hi = tb.tw_hi;
  a.scale_a = scale_a; a.scale_b = scale_b; a.scale_split = scale_split;
  a.K = K; a.logS = logS; a.logN = logN; a.kind = kind; a.inverse = inverse ? 1 : 0;
  const int log_groups = logN - K;        // number of (sub, t) pairs per column
  a.contig = (logS == 0) ? 1 : 0;
  int logT = MAX_TILE_LOG - K;
  if (logT < 0) logT = 0;
  if (a.contig) { if (logT > log_groups) logT = log_groups; if (kind == PASS_DIT_FINAL) logT = 0; }
  else if (logT > logS) logT = logS;
  a.logT = logT;
  size_t row = a.contig ? ((1u << K) + 1) : ((1u << logT) + 1);
  size_t rows = a.contig ? (1u << logT) : (1u << K);
  size_t smem = (((1u << K) >> 1) + 1 + row * rows) * sizeof(u32);
  dim3 grid(1u << (log_groups - logT), (unsigned)ncols);
  ntt_pass_kernel<<<grid, NTT_THREADS, smem, s>>>(a);
  ZKB_CHECK_LAUNCH();
}

// split logN into pass sizes, each <= 8 bits, as evenly as possible (first entries larger)
static std::vector<int> plan_passes(int logN) {
  std::vector<int> ks;
  if (logN == 0) return ks;
  int np = (logN + 7) / 8;
  for (int i = 0; i < np; i++) ks.push_back(logN / np + (i < logN % np ? 1 : 0));
  return ks;
}

// DIF transform (natural -> bit-reversed) of ncols columns; first pass in -> out, rest in place
static void run_dif(const NttTables& tb, const u32* in, size_t in_stride, u32* out, size_t out_stride, size_t ncols,
                    int logN, bool inverse, cudaStream_t s) {
  std::vector<int> ks = plan_passes(logN);
  int logL = logN;
  for (size_t p = 0; p < ks.size(); p++) {
    int K = ks[p], logS = logL - K;
    launch_pass(tb, p == 0 ? in : out, p == 0 ? in_stride : out_stride, out, out_stride, ncols, logN, K, logS, PASS_DIF,
                inverse, nullptr, nullptr, 0, s);
    logL = logS;
  }
}

// DIT transform (bit-reversed -> bit-reversed store) with optional load scaling:
// in -> (scratch, in place) -> out.  `scratch` must hold ncols columns (stride scratch_stride).
static void run_dit_bitrev_out(const NttTables& tb, const u32* in, size_t in_stride, u32* scratch, size_t scratch_stride,
                               u32* out, size_t out_stride, size_t ncols, int logN, bool inverse, const u32* scale_a,
                               const u32* scale_b, int scale_split, cudaStream_t s) {
  std::vector<int> ks = plan_passes(logN);
  // DIT runs the levels bottom-up: contiguous groups first, the stride-N/2^K level last
  std::reverse(ks.begin(), ks.end());
  int logS = 0;
  for (size_t p = 0; p < ks.size(); p++) {
    int K = ks[p];
    bool last = (p + 1 == ks.size());
    const u32* src = p == 0 ? in : scratch;
    size_t src_stride = p == 0 ? in_stride : scratch_stride;
    launch_pass(tb, src, src_stride, last ? out : scratch, last ? out_stride : scratch_stride, ncols, logN, K, logS,
                last ? PASS_DIT_FINAL : PASS_DIT, inverse, p == 0 ? scale_a : nullptr, p == 0 ? scale_b : nullptr,
                scale_split, s);
    logS += K;
  }
}

// scale tables for "coefficient at bit-reversed position p gets c * shift^bitrev(p)"
__global__ void scale_tables_kernel(u32* a, u32* b, int logn, int split, Fp shift, Fp c) {
  u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  u32 na = 1u << split, nb = 1u << (logn - split);
  if (i < na) a[i] = (c * fp_pow(shift, (u64)bitrev32(i, split) << (logn - split))).v;
  if (i < nb) b[i] = fp_pow(shift, (u64)bitrev32(i, logn - split)).v;
}

void coset_lde_batch(const NttTables& tb, const u32* in, size_t in_stride, u32* out, size_t out_stride,
                     unsigned log_n, size_t width, unsigned log_blowup, Fp shift, cudaStream_t s) {
  if (width == 0) return;
  const size_t n = (size_t)1 << log_n;
  const unsigned ncoset = 1u << log_blowup;
  const int split = (int)(log_n + 1) / 2;
  // per-coset scale tables
  DevBuf tabs(((size_t)ncoset) * (((size_t)1 << split) + ((size_t)1 << (log_n - split))), s);
  const size_t tab_sz = ((size_t)1 << split) + ((size_t)1 << (log_n - split));
  Fp ninv = fp_inv(fp_from_canonical((u32)(n % KB_P)));
  Fp wN = two_adic_generator(log_n + log_blowup);
  for (unsigned c = 0; c < ncoset; c++) {
    Fp sh = shift * fp_pow(wN, c);
    u32* ta = tabs.p + c * tab_sz;
    u32* tbp = ta + ((size_t)1 << split);
    unsigned cnt = 1u << (split > (int)log_n - split ? split : (int)log_n - split);
    scale_tables_kernel<<<ceil_div(cnt, 256), 256, 0, s>>>(ta, tbp, (int)log_n, split, sh, ninv);
    ZKB_CHECK_LAUNCH();
  }
  // column chunks of ~32 MB so that pass-to-pass traffic stays in L2
  size_t chunk = ((size_t)1 << 23) >> log_n;
  if (chunk < 1) chunk = 1;
  if (chunk > width) chunk = width;
  if (chunk > 32768) chunk = 32768;
  DevBuf coef(chunk * n, s), scratch(chunk * n, s);
  for (size_t c0 = 0; c0 < width; c0 += chunk) {
    size_t nc = width - c0 < chunk ? width - c0 : chunk;
    if (log_n == 0) {
      // constant columns: every coset evaluation equals the single value
      for (unsigned c = 0; c < ncoset; c++)
        ZKB_CUDA(cudaMemcpy2DAsync(out + c0 * out_stride + c, out_stride * 4, in + c0 * in_stride, in_stride * 4, 4, nc,
                                   cudaMemcpyDeviceToDevice, s));
      continue;
    }
    run_dif(tb, in + c0 * in_stride, in_stride, coef.p, n, nc, (int)log_n, true, s);
    for (unsigned c = 0; c < ncoset; c++) {
      u32* ta = tabs.p + c * tab_sz;
      u32* tbp = ta + ((size_t)1 << split);
      // coset c (shift * w_N^c) lives in block bitrev(c) of the bit-reversed output
      size_t blk = bitrev32(c, log_blowup);
      run_dit_bitrev_out(tb, coef.p, n, scratch.p, n, out + c0 * out_stride + blk * n, out_stride, nc, (int)log_n, false,
                         ta, tbp, split, s);
    }
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
  for (size_t c0 = 0; c0 < width; c0 += 32768) {
    size_t nc = width - c0 < 32768 ? width - c0 : 32768;
    if (bitrev_out) {
      run_dif(tb, in + c0 * n, n, out + c0 * n, n, nc, (int)log_n, inverse, s);
    } else {
      DevBuf tmp(nc * n, s);
      run_dif(tb, in + c0 * n, n, tmp.p, n, nc, (int)log_n, inverse, s);
      bitrev_rows(tmp.p, out + c0 * n, log_n, nc, s);
    }
  }
  if (inverse) {
    Fp ninv = fp_inv(fp_from_canonical((u32)(n % KB_P)));
    size_t count = n * width;
    scale_all_kernel<<<ceil_div(count, 256), 256, 0, s>>>(out, count, ninv);
    ZKB_CHECK_LAUNCH();
  }
}

}  // namespace zkb

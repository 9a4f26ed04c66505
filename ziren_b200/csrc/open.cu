#include "open.h"
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>

namespace zkb {

std::atomic<int> g_eval_v2{-1};    // zkb200_set_option("eval_v2"): -1 = environment / default

__global__ void bary_weights_kernel(const u32* tw_lo, const u32* tw_hi, unsigned log_n, Ef zp, Ef scale, u32* out) {
  const size_t n = (size_t)1 << log_n;
  size_t pos = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (pos >= n) return;
  u32 i = bitrev32((u32)pos, log_n);
  Fp h = tw_pow2(tw_lo, tw_hi, i << (24 - log_n));
  Ef w = (ef_inv(zp - h) * h) * scale;
#pragma unroll
  for (int c = 0; c < 4; c++) out[(size_t)c * n + pos] = w.c[c].v;
}
void bary_weights(const NttTables& tb, unsigned log_n, const Ef& z, u32* out, cudaStream_t s) {
  const size_t n = (size_t)1 << log_n;
  // z' = z / g;  scale = (z'^n - 1) / n
  Ef zp = z * fp_inv(fp_from_canonical(KB_GEN));
  Ef zn = zp;
  for (unsigned i = 0; i < log_n; i++) zn *= zn;
  Ef scale = (zn - fp_one()) * fp_inv(fp_from_canonical((u32)(n % KB_P)));
  bary_weights_kernel<<<ceil_div(n, 128), 128, 0, s>>>(tb.tw_lo, tb.tw_hi, log_n, zp, scale, out);
  ZKB_CHECK_LAUNCH();
}

// ---- column evaluation --------------------------------------------------------------------------
constexpr int EC_COLS = 4;       // columns per CTA
constexpr int EC_THREADS = 256;

template <int NPT>
__global__ void __launch_bounds__(EC_THREADS) eval_columns_kernel(const u32* __restrict__ lde, size_t H, size_t n, size_t W,
                                                                  const u32* __restrict__ w0, const u32* __restrict__ w1,
                                                                  u32* __restrict__ partial, size_t rows_per_split) {
  const size_t c0 = (size_t)blockIdx.x * EC_COLS;
  const size_t r_begin = (size_t)blockIdx.y * rows_per_split;
  size_t r_end = r_begin + rows_per_split;
  if (r_end > n) r_end = n;
  Ef acc[NPT][EC_COLS];
#pragma unroll
  for (int p = 0; p < NPT; p++)
#pragma unroll
    for (int c = 0; c < EC_COLS; c++) acc[p][c] = ef_zero();
  // four rows per iteration: the four products of a component are summed raw in 64 bits
  // (4 p^2 < 2^64) and reduced once
  for (size_t rb = r_begin + threadIdx.x; rb < r_end; rb += 4 * EC_THREADS) {
    u32 wv[4][NPT][4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const size_t r = rb + (size_t)u * EC_THREADS;
      const bool ok = r < r_end;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        wv[u][0][k] = ok ? w0[(size_t)k * n + r] : 0;
        if (NPT > 1) wv[u][NPT - 1][k] = ok ? w1[(size_t)k * n + r] : 0;
      }
    }
#pragma unroll
    for (int c = 0; c < EC_COLS; c++) {
      if (c0 + c < W) {
        u32 x[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const size_t r = rb + (size_t)u * EC_THREADS;
          x[u] = r < r_end ? lde[(c0 + c) * H + r] : 0;
        }
#pragma unroll
        for (int p = 0; p < NPT; p++)
#pragma unroll
          for (int k = 0; k < 4; k++) {
            u64 s = (u64)wv[0][p][k] * x[0] + (u64)wv[1][p][k] * x[1] + (u64)wv[2][p][k] * x[2] + (u64)wv[3][p][k] * x[3];
            acc[p][c].c[k] += fp_raw(mont_reduce_wide(s));
          }
      }
    }
  }
  // block reduction: warp shuffle, then across warps through shared memory
  __shared__ u32 sh[EC_THREADS / 32][NPT * EC_COLS * 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int p = 0; p < NPT; p++)
#pragma unroll
    for (int c = 0; c < EC_COLS; c++)
#pragma unroll
      for (int k = 0; k < 4; k++) {
        Fp v = acc[p][c].c[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = v + fp_raw(__shfl_xor_sync(0xffffffffu, v.v, d));
        if (lane == 0) sh[warp][(p * EC_COLS + c) * 4 + k] = v.v;
      }
  __syncthreads();
  for (int idx = threadIdx.x; idx < NPT * EC_COLS * 4; idx += EC_THREADS) {
    Fp v = fp_zero();
    for (int wv = 0; wv < EC_THREADS / 32; wv++) v = v + fp_raw(sh[wv][idx]);
    int p = idx / (EC_COLS * 4), c = (idx / 4) % EC_COLS, k = idx % 4;
    if (c0 + c < W) partial[(((size_t)blockIdx.y * NPT + p) * W + c0 + c) * 4 + k] = v.v;
  }
}
// ---- column evaluation, variant 2 (opt-in: ZKB200_EVAL_V2=1; not yet measured on a B200) --------
// One LANE per column instead of one thread per row: a warp stages a 32-column x 32-row tile of the
// matrix in shared memory with coalesced loads (cp.async, double-buffered), then every lane walks ITS
// column of the tile while the barycentric weights of the row are broadcast from shared memory.  The
// accumulators of a column live in one thread for the whole row range, so there is no cross-thread
// reduction, the weights are read once per CTA and step instead of once per four columns, and the
// kernel needs about 50 registers (the row-parallel kernel above: 128 registers, 12 % occupancy,
// 830 GB/s on the 2^18 x 4167 matrix; the multiply work alone allows about 2.4x that).
constexpr int E2_WARPS = 8;      // 8 warps = 256 columns per CTA
constexpr int E2_ROWS = 32;      // rows per step
constexpr int E2_LD = 33;        // tile row stride: lane-along-column reads are conflict-free

__device__ __forceinline__ void cp_async_4(u32* smem_dst, const u32* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int NPT>
__global__ void __launch_bounds__(E2_WARPS * 32, 3) eval_columns_v2_kernel(const u32* __restrict__ lde, size_t H, size_t n, size_t W,
                                                                           const u32* __restrict__ w0, const u32* __restrict__ w1,
                                                                           u32* __restrict__ partial, size_t rows_per_split) {
  extern __shared__ __align__(16) u32 e2_smem[];
  // [2 buffers][E2_WARPS][32 columns][E2_LD] data words, then [2 buffers][E2_ROWS][8] weight words
  u32* tiles = e2_smem;
  u32* wts = e2_smem + 2 * E2_WARPS * 32 * E2_LD;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t c0 = ((size_t)blockIdx.x * E2_WARPS + warp) * 32;     // first column of this warp
  const size_t col = c0 + lane;
  const size_t r_begin = (size_t)blockIdx.y * rows_per_split;        // multiples of E2_ROWS (host side)
  size_t r_end = r_begin + rows_per_split;
  if (r_end > n) r_end = n;
  const u32 steps = r_begin < r_end ? (u32)((r_end - r_begin) / E2_ROWS) : 0;

  // stage step `st` into buffer `buf`: lane = row within the step, one 128-byte segment per column
  auto stage = [&](u32 st, int buf) {
    const size_t r0 = r_begin + (size_t)st * E2_ROWS;
    u32* t = tiles + ((size_t)buf * E2_WARPS + warp) * 32 * E2_LD;
#pragma unroll 8
    for (int j = 0; j < 32; j++)
      if (c0 + j < W) cp_async_4(t + j * E2_LD + lane, lde + (c0 + j) * H + r0 + lane);
    // weights of the step: component k of row rr at wts[buf][rr][k]; threads run along rows (coalesced)
    if (threadIdx.x < E2_ROWS * NPT * 4) {
      const int k = threadIdx.x / E2_ROWS, rr = threadIdx.x % E2_ROWS;
      const u32* src = (k < 4 ? w0 + (size_t)k * n : w1 + (size_t)(k - 4) * n) + r0 + rr;
      cp_async_4(wts + ((size_t)buf * E2_ROWS + rr) * 8 + k, src);
    }
    cp_async_commit();
  };

  Fp tot[NPT][4];
  u64 raw[NPT][4];
#pragma unroll
  for (int p = 0; p < NPT; p++)
#pragma unroll
    for (int k = 0; k < 4; k++) { tot[p][k] = fp_zero(); raw[p][k] = 0; }

  if (steps) stage(0, 0);
  for (u32 st = 0; st < steps; st++) {
    const int buf = st & 1;
    cp_async_wait_all();
    __syncthreads();                       // step st is in shared memory; buffer buf^1 is free (see below)
    if (st + 1 < steps) stage(st + 1, buf ^ 1);
    const u32* t = tiles + ((size_t)buf * E2_WARPS + warp) * 32 * E2_LD + lane * E2_LD;   // this lane's column
    const u32* wr = wts + (size_t)buf * E2_ROWS * 8;
#pragma unroll 2
    for (int r4 = 0; r4 < E2_ROWS; r4 += 4) {
      // four rows: the four products of a component are summed raw in 64 bits (4 p^2 < 2^64)
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const u32 x = t[r4 + u];
        const uint4 a = *reinterpret_cast<const uint4*>(wr + (r4 + u) * 8);
        raw[0][0] += (u64)a.x * x; raw[0][1] += (u64)a.y * x; raw[0][2] += (u64)a.z * x; raw[0][3] += (u64)a.w * x;
        if (NPT > 1) {
          const uint4 b = *reinterpret_cast<const uint4*>(wr + (r4 + u) * 8 + 4);
          raw[NPT - 1][0] += (u64)b.x * x; raw[NPT - 1][1] += (u64)b.y * x;
          raw[NPT - 1][2] += (u64)b.z * x; raw[NPT - 1][3] += (u64)b.w * x;
        }
      }
#pragma unroll
      for (int p = 0; p < NPT; p++)
#pragma unroll
        for (int k = 0; k < 4; k++) { tot[p][k] = tot[p][k] + fp_raw(mont_reduce_wide(raw[p][k])); raw[p][k] = 0; }
    }
    // the next iteration's __syncthreads (after its wait) orders these reads before buffer `buf` is
    // staged again two steps later
  }
  if (col < W) {
#pragma unroll
    for (int p = 0; p < NPT; p++)
#pragma unroll
      for (int k = 0; k < 4; k++) partial[(((size_t)blockIdx.y * NPT + p) * W + col) * 4 + k] = tot[p][k].v;
  }
}

// SM count, occupancy and the opt-in shared-memory size are per-device facts: cached per device
// (a process may hold contexts on several GPUs), filled under a lock on first use.
struct OpenDeviceInfo { bool ready = false; int sms = 148; int slots_per_sm[2] = {1, 1}; };
static const OpenDeviceInfo& open_device_info() {
  static std::mutex mu;
  static OpenDeviceInfo info[64];
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  OpenDeviceInfo& d = info[dev & 63];
  if (!d.ready) {
    cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.slots_per_sm[0], eval_columns_kernel<1>, EC_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d.slots_per_sm[1], eval_columns_kernel<2>, EC_THREADS, 0);
    const size_t smem = (size_t)(2 * E2_WARPS * 32 * E2_LD + 2 * E2_ROWS * 8) * sizeof(u32);
    cudaFuncSetAttribute(eval_columns_v2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(eval_columns_v2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (d.sms <= 0) d.sms = 148;
    d.ready = true;
  }
  return d;
}

__global__ void sum_partials_kernel(const u32* partial, size_t count, int nsplit, u32* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  Fp v = fp_zero();
  for (int s = 0; s < nsplit; s++) v = v + fp_raw(partial[(size_t)s * count + i]);
  out[i] = v.v;
}
void eval_columns(const u32* lde, size_t H, size_t n, size_t W, const u32* w0, const u32* w1, int npoints, u32* out,
                  cudaStream_t s) {
  if (!W) return;
  static const bool env_v2 = getenv("ZKB200_EVAL_V2") && atoi(getenv("ZKB200_EVAL_V2")) != 0;
  static const bool env_set = getenv("ZKB200_EVAL_V2") != nullptr;
  // default ON: measured 10.5 ms against 12.0 ms for the opening stage of the bench shard (profiles/README.md)
  const bool use_v2 = g_eval_v2.load() >= 0 ? g_eval_v2.load() != 0 : (env_set ? env_v2 : true);
  if (use_v2 && n >= 1024 && W >= 128) {
    // rows are split in multiples of E2_ROWS so that every CTA runs whole steps; about three CTAs per SM
    const size_t smem = (size_t)(2 * E2_WARPS * 32 * E2_LD + 2 * E2_ROWS * 8) * sizeof(u32);
    const int v2_sms = open_device_info().sms;
    const size_t colgroups = (W + E2_WARPS * 32 - 1) / (E2_WARPS * 32);
    // ONE full wave of equal CTAs (3 per SM): rounding the split count UP put 459 CTAs on 444 slots, i.e. a
    // second wave of 15 CTAs that doubled the kernel's time
    size_t nsplit = ((size_t)3 * v2_sms) / colgroups;
    if (nsplit > n / 256) nsplit = n / 256;
    if (nsplit < 1) nsplit = 1;
    size_t rows_per_split = ((n + nsplit - 1) / nsplit + E2_ROWS - 1) / E2_ROWS * E2_ROWS;
    nsplit = (n + rows_per_split - 1) / rows_per_split;
    const size_t count = (size_t)npoints * W * 4;
    DevBuf partial(nsplit > 1 ? nsplit * count : 0, s);
    u32* dst = nsplit > 1 ? partial.p : out;
    dim3 grid((unsigned)colgroups, (unsigned)nsplit);
    if (npoints == 1) eval_columns_v2_kernel<1><<<grid, E2_WARPS * 32, smem, s>>>(lde, H, n, W, w0, w0, dst, rows_per_split);
    else eval_columns_v2_kernel<2><<<grid, E2_WARPS * 32, smem, s>>>(lde, H, n, W, w0, w1, dst, rows_per_split);
    ZKB_CHECK_LAUNCH();
    if (nsplit > 1) {
      sum_partials_kernel<<<ceil_div(count, 256), 256, 0, s>>>(partial.p, count, (int)nsplit, out);
      ZKB_CHECK_LAUNCH();
    }
    return;
  }
  const unsigned colblocks = ceil_div(W, EC_COLS);
  // Row splits: enough CTAs to fill the GPU, and for large grids the split count (<= 8) whose last
  // wave is fullest (every CTA does the same work, so a ragged last wave is lost time: the
  // 1042-CTA keccak matrix runs 3.5 waves unsplit)
  const OpenDeviceInfo& di = open_device_info();
  const int sms = di.sms;
  const int* slots_per_sm = di.slots_per_sm;
  const size_t slots = (size_t)sms * std::max(1, slots_per_sm[npoints == 1 ? 0 : 1]);
  size_t nsplit = 1;
  if (colblocks < slots) nsplit = (slots + colblocks - 1) / colblocks;
  else {
    double best = 0;
    for (size_t k = 1; k <= 8; k++) {
      const double waves = (double)(colblocks * k) / (double)slots;
      const double eff = waves / (double)(size_t)(waves + 0.999999);
      if (eff > best + 0.02) { best = eff; nsplit = k; }
    }
  }
  size_t max_split = n / 1024 ? n / 1024 : 1;
  if (nsplit > max_split) nsplit = max_split;
  size_t rows_per_split = (n + nsplit - 1) / nsplit;
  const size_t count = (size_t)npoints * W * 4;
  DevBuf partial(nsplit > 1 ? nsplit * count : 0, s);
  u32* dst = nsplit > 1 ? partial.p : out;
  dim3 grid(colblocks, (unsigned)nsplit);
  if (npoints == 1) eval_columns_kernel<1><<<grid, EC_THREADS, 0, s>>>(lde, H, n, W, w0, w0, dst, rows_per_split);
  else eval_columns_kernel<2><<<grid, EC_THREADS, 0, s>>>(lde, H, n, W, w0, w1, dst, rows_per_split);
  ZKB_CHECK_LAUNCH();
  if (nsplit > 1) {
    sum_partials_kernel<<<ceil_div(count, 256), 256, 0, s>>>(partial.p, count, (int)nsplit, out);
    ZKB_CHECK_LAUNCH();
  }
}

__global__ void inv_denoms_kernel(const u32* tw_lo, const u32* tw_hi, unsigned log_h, Ef z, Fp gen, u32* out) {
  const size_t H = (size_t)1 << log_h;
  size_t pos = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (pos >= H) return;
  u32 i = bitrev32((u32)pos, log_h);
  Fp x = gen * tw_pow2(tw_lo, tw_hi, i << (24 - log_h));
  Ef d = ef_inv(z - x);
#pragma unroll
  for (int c = 0; c < 4; c++) out[(size_t)c * H + pos] = d.c[c].v;
}
void inv_denominators(const NttTables& tb, unsigned log_h, const Ef& z, u32* out, cudaStream_t s) {
  const size_t H = (size_t)1 << log_h;
  inv_denoms_kernel<<<ceil_div(H, 128), 128, 0, s>>>(tb.tw_lo, tb.tw_hi, log_h, z, fp_from_canonical(KB_GEN), out);
  ZKB_CHECK_LAUNCH();
}

__global__ void ef_powers_kernel(Ef alpha, size_t count, u32* out) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= count) return;
  Ef p = ef_pow(alpha, j);
  uint4 v = make_uint4(p.c[0].v, p.c[1].v, p.c[2].v, p.c[3].v);
  reinterpret_cast<uint4*>(out)[j] = v;
}
void ef_powers(const Ef& alpha, size_t count, u32* out, cudaStream_t s) {
  if (!count) return;
  ef_powers_kernel<<<ceil_div(count, 128), 128, 0, s>>>(alpha, count, out);
  ZKB_CHECK_LAUNCH();
}

// ---- reduced openings ---------------------------------------------------------------------------
__device__ __forceinline__ Ef ldg_ef(const u32* p, size_t j) {
  uint4 v = __ldg(reinterpret_cast<const uint4*>(p) + j);
  Ef e; e.c[0] = fp_raw(v.x); e.c[1] = fp_raw(v.y); e.c[2] = fp_raw(v.z); e.c[3] = fp_raw(v.w);
  return e;
}
// rys[pt] = sum_j alpha^j ys[pt][j]   (one CTA per point)
__global__ void __launch_bounds__(256) reduce_ys_kernel(const u32* ys, const u32* apow, size_t W, u32* rys) {
  const u32* y = ys + (size_t)blockIdx.x * W * 4;
  Ef acc = ef_zero();
  for (size_t j = threadIdx.x; j < W; j += blockDim.x) acc += ldg_ef(apow, j) * ldg_ef(y, j);
  __shared__ u32 sh[8][4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    Fp v = acc.c[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = v + fp_raw(__shfl_xor_sync(0xffffffffu, v.v, d));
    if (lane == 0) sh[warp][k] = v.v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    Fp v = fp_zero();
    for (int w = 0; w < 8; w++) v = v + fp_raw(sh[w][threadIdx.x]);
    rys[blockIdx.x * 4 + threadIdx.x] = v.v;
  }
}
// Two rows per thread (rows x and x + 128 of a 256-row block): the alpha powers, which every thread
// loads per column, are used twice, and 16 column loads are in flight per thread.
std::atomic<int> g_k4b_rows{2};     // zkb200_set_option("k4b_rows"): 1, 2 or 4 rows per thread
template <int NPT, int RM_ROWS>
__global__ void __launch_bounds__(128) reduce_matrix_kernel(const u32* __restrict__ lde, size_t H, size_t W,
                                                            const u32* __restrict__ apow, const u32* __restrict__ rys,
                                                            Ef off0, Ef off1, const u32* __restrict__ invden0,
                                                            const u32* __restrict__ invden1, u32* __restrict__ ro) {
  const size_t x0 = (size_t)blockIdx.x * (128 * RM_ROWS) + threadIdx.x;
  // rows past the end read row H - 1 again (results discarded), so the loop body is branch free
  size_t xs[RM_ROWS];
#pragma unroll
  for (int r = 0; r < RM_ROWS; r++) { xs[r] = x0 + (size_t)r * 128; if (xs[r] >= H) xs[r] = H - 1; }
  // eight columns per iteration: the loads are independent and issued together, and every
  // group of four products is reduced once (4 p^2 < 2^64)
  Ef acc[RM_ROWS];
#pragma unroll
  for (int r = 0; r < RM_ROWS; r++) acc[r] = ef_zero();
  size_t j = 0;
  for (; j + 8 <= W; j += 8) {
    u32 v[RM_ROWS][8];
#pragma unroll
    for (int u = 0; u < 8; u++)
#pragma unroll
      for (int r = 0; r < RM_ROWS; r++) v[r][u] = lde[(j + u) * H + xs[r]];
#pragma unroll
    for (int g = 0; g < 2; g++) {
      u64 s[RM_ROWS][4];
#pragma unroll
      for (int r = 0; r < RM_ROWS; r++)
#pragma unroll
        for (int k = 0; k < 4; k++) s[r][k] = 0;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const Ef a = ldg_ef(apow, j + 4 * g + u);
#pragma unroll
        for (int r = 0; r < RM_ROWS; r++)
#pragma unroll
          for (int k = 0; k < 4; k++) s[r][k] += (u64)a.c[k].v * v[r][4 * g + u];
      }
#pragma unroll
      for (int r = 0; r < RM_ROWS; r++)
#pragma unroll
        for (int k = 0; k < 4; k++) acc[r].c[k] += fp_raw(mont_reduce_wide(s[r][k]));
    }
  }
  if (j < W) {
    EfAcc lazy[RM_ROWS];
#pragma unroll
    for (int r = 0; r < RM_ROWS; r++) lazy[r].clear();
    for (; j < W; j++) {
      const Ef a = ldg_ef(apow, j);
#pragma unroll
      for (int r = 0; r < RM_ROWS; r++) lazy[r].add(a, fp_raw(lde[j * H + xs[r]]));
    }
#pragma unroll
    for (int r = 0; r < RM_ROWS; r++) acc[r] += lazy[r].value();
  }
#pragma unroll
  for (int r = 0; r < RM_ROWS; r++) {
    const size_t x = x0 + (size_t)r * 128;
    if (x >= H) continue;
    Ef res;
#pragma unroll
    for (int c = 0; c < 4; c++) res.c[c] = fp_raw(ro[(size_t)c * H + x]);
    {
      Ef d;
#pragma unroll
      for (int c = 0; c < 4; c++) d.c[c] = fp_raw(invden0[(size_t)c * H + x]);
      res += off0 * ((ldg_ef(rys, 0) - acc[r]) * d);
    }
    if (NPT > 1) {
      Ef d;
#pragma unroll
      for (int c = 0; c < 4; c++) d.c[c] = fp_raw(invden1[(size_t)c * H + x]);
      res += off1 * ((ldg_ef(rys, 1) - acc[r]) * d);
    }
#pragma unroll
    for (int c = 0; c < 4; c++) ro[(size_t)c * H + x] = res.c[c].v;
  }
}
void reduce_matrix(const u32* lde, size_t H, size_t W, const u32* apow, const u32* ys, int npoints, const Ef& off0,
                   const Ef& off1, const u32* invden0, const u32* invden1, u32* ro, cudaStream_t s) {
  if (!W) return;
  DevBuf rys(8, s);
  reduce_ys_kernel<<<npoints, 256, 0, s>>>(ys, apow, W, rys.p);
  ZKB_CHECK_LAUNCH();
  const int rows = g_k4b_rows.load();
#define ZKB_RM(NPT, R, D1) reduce_matrix_kernel<NPT, R><<<ceil_div(H, 128 * R), 128, 0, s>>>(lde, H, W, apow, rys.p, off0, off1, invden0, D1, ro)
  if (npoints == 1) { if (rows >= 4) ZKB_RM(1, 4, invden0); else if (rows == 2) ZKB_RM(1, 2, invden0); else ZKB_RM(1, 1, invden0); }
  else { if (rows >= 4) ZKB_RM(2, 4, invden1); else if (rows == 2) ZKB_RM(2, 2, invden1); else ZKB_RM(2, 1, invden1); }
#undef ZKB_RM
  ZKB_CHECK_LAUNCH();
}

}  // namespace zkb

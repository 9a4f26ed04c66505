// K1, "lean" strided level: the same four-step strided NTT level as ntt_strided_kernel (ntt.cu), with
// EVERY size a compile-time constant (level size 2^K, stride 2^LOGS, tile width 32 = one warp), so
// that all shared-memory, twiddle and global addresses are one per-thread base register plus an
// immediate.  The generic kernel spends two thirds of its instructions on address arithmetic
// (SASS: 1250 instructions per 16 elements, 450 of them butterflies); this one has none in its body:
//   * tile [2^K slots][32 offsets], plain row-major: a warp always touches one 128-byte row, so there
//     is no swizzle and no bank conflict by construction;
//   * thread = (warp w, lane q): its 16 elements of a radix round are slots ins(w) | const, i.e.
//     base + immediate, for loads, stores and the global rows alike;
//   * butterfly twiddles are warp-uniform (they depend on w only): one broadcast LDS.64 each, no padding;
//   * the final DIT level leaves through a [32][2^K + 1] staging tile, so the bit-reversed rows are
//     written as 128-byte runs of 2^K-word rows.
// Included by ntt.cu (uses its shoup_mul / round_bits / round_lo helpers).
#pragma once
#include <type_traits>

namespace zkb {

struct LeanStridedArgs {
  const u32* in; u32* out;
  size_t in_stride, out_stride;          // elements between columns
  size_t in_coset_stride, out_coset_stride;
  const uint2* small_tw;                 // Shoup pairs of w_{2^K}^(+-e) (NttTables::small_tw)
  const uint2* four;                     // four-step twiddles by element position
  int inverse;
};

// insert R zero bits at bit LO
template <int R, int LO>
__host__ __device__ constexpr u32 lean_ins(u32 rest) { return ((rest >> LO) << (LO + R)) | (rest & ((1u << LO) - 1u)); }
__host__ __device__ constexpr u32 lean_bitrev(u32 x, int bits) {
  u32 r = 0;
  for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}

// radix-2^R DIF butterflies on x[0 .. 2^R): x[m] is slot (base | m << LO); the twiddle of the pair
// (m0, m1) at level lh is w^((bl << sh) + (k << (LO + sh))), bl = base & (2^LO - 1) = blw | BLC.
template <int K, int R, int LO, u32 BLC>
__device__ __forceinline__ void lean_butterflies(Fp* x, u32 blw, const uint2* __restrict__ tw) {
#pragma unroll
  for (int j = 0; j < R; j++) {
    const int lh = R - 1 - j;
    const int sh = K - 1 - LO - lh;
    const uint2* __restrict__ twl = tw + (blw << sh);
#pragma unroll
    for (int p = 0; p < (1 << (R - 1)); p++) {
      const int k = p & ((1 << lh) - 1);
      const int m0 = ((p >> lh) << (lh + 1)) | k;
      const int m1 = m0 + (1 << lh);
      const uint2 w = twl[(BLC << sh) + ((u32)k << (LO + sh))];
      const Fp a = x[m0], b = x[m1];
      x[m0] = a + b;
      x[m1] = shoup_mul(a.v - b.v + KB_P, w);
    }
  }
}

// compile-time loop: f(integral_constant<int, 0>) ... f(integral_constant<int, N-1>)
template <int I, int N, class F>
__device__ __forceinline__ void lean_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>());
    lean_for<I + 1, N>(f);
  }
}

// One radix round over bits [LO, LO + R) of the slot index: 16 elements per thread in G groups.
// load(cg, dst) / store(cg, v): cg is the COMPILE-TIME part of the slot (group and butterfly position),
// the per-thread part (warp w, lane) is folded into the callers' base pointers.
template <int K, int R, int LO, int NT, class Load, class Store>
__device__ __forceinline__ void lean_round(u32 w, const uint2* __restrict__ tw, Load load, Store store) {
  constexpr int G = 16 >> R;
  constexpr u32 NW = NT / 32;
  const u32 blw = lean_ins<R, LO>(w) & ((1u << LO) - 1u);
  Fp x[G][1 << R];
  lean_for<0, G>([&](auto gc) {
    constexpr int g = decltype(gc)::value;
    lean_for<0, (1 << R)>([&](auto mc) {
      constexpr int m = decltype(mc)::value;
      x[g][m] = load(std::integral_constant<u32, (lean_ins<R, LO>(g * NW) | ((u32)m << LO))>());
    });
  });
  lean_for<0, G>([&](auto gc) {
    constexpr int g = decltype(gc)::value;
    lean_butterflies<K, R, LO, (lean_ins<R, LO>(g * NW) & ((1u << LO) - 1u))>(x[g], blw, tw);
    lean_for<0, (1 << R)>([&](auto mc) {
      constexpr int m = decltype(mc)::value;
      store(std::integral_constant<u32, (lean_ins<R, LO>(g * NW) | ((u32)m << LO))>(), x[g][m]);
    });
  });
}

template <int K>
__host__ __device__ constexpr size_t lean_strided_smem(bool final_dit) {
  return (((size_t)1 << (K - 1)) * 2 + ((size_t)32 << K) + (final_dit ? (size_t)32 * ((1u << K) + 1) : 0)) * sizeof(u32);
}

template <int K, int LOGS, bool FINAL>
__global__ void __launch_bounds__(1 << (K + 1), (K <= 8 ? 2 : 1)) ntt_strided_lean_kernel(LeanStridedArgs a) {
  static_assert(K >= 4 && K <= 9 && LOGS >= 5, "lean strided level: 2^(K+1) threads, 32-offset tiles");
  constexpr int NT = 1 << (K + 1);
  constexpr u32 NW = NT / 32;
  constexpr int NR = (K + 3) / 4;
  constexpr u32 P = (1u << K) + 1;             // row pitch of the final staging tile
  extern __shared__ __align__(16) u32 smem[];
  uint2* tw = reinterpret_cast<uint2*>(smem);                    // 2^(K-1) pairs
  u32* tile = smem + (1u << K);                                  // [2^K][32]
  u32* stage = tile + (32u << K);                                // FINAL: [32][P]
  const u32 lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const u32 t0 = blockIdx.x << 5;
  const u32* __restrict__ in = a.in + (size_t)blockIdx.y * a.in_stride + (size_t)blockIdx.z * a.in_coset_stride;
  u32* __restrict__ out = a.out + (size_t)blockIdx.y * a.out_stride + (size_t)blockIdx.z * a.out_coset_stride;
  {
    constexpr u32 half = 1u << (K - 1);
    const uint2* src = a.small_tw + (a.inverse ? ((1u << 12) - 1) : 0) + (half - 1);      // KMAX_TAB = 12
    for (u32 m = threadIdx.x; m < half; m += NT) tw[m] = src[m];
  }
  __syncthreads();

  // ---- round 0: straight from global memory ----
  constexpr int R0 = round_bits(K, 0), LO0 = round_lo(K, 0);
  {
    const u32 sw = lean_ins<R0, LO0>(w);                         // per-thread part of the slot
    const u32 roww = FINAL ? (__brev(sw) >> (32 - K)) : sw;
    const u32* __restrict__ gp = in + ((size_t)roww << LOGS) + t0 + lane;
    const uint2* __restrict__ fp4 = a.four + ((size_t)roww << LOGS) + t0 + lane;
    u32* __restrict__ sp = tile + sw * 32 + lane;
    auto gload = [&](auto cg) {
      constexpr u32 rowc = FINAL ? lean_bitrev(decltype(cg)::value, K) : decltype(cg)::value;
      const u32 v = gp[(size_t)rowc << LOGS];
      if constexpr (FINAL) return shoup_mul(v, __ldg(fp4 + ((size_t)rowc << LOGS)));
      else return fp_raw(v);
    };
    if constexpr (NR == 1) {
      if constexpr (FINAL) {
        u32* __restrict__ st = stage + lane * P + sw;
        lean_round<K, R0, LO0, NT>(w, tw, gload, [&](auto cg, Fp v) { st[decltype(cg)::value] = v.v; });
      } else {
        u32* __restrict__ op = out + ((size_t)sw << LOGS) + t0 + lane;
        const uint2* __restrict__ f2 = a.four + ((size_t)sw << LOGS) + t0 + lane;
        lean_round<K, R0, LO0, NT>(w, tw, gload, [&](auto cg, Fp v) {
          constexpr size_t off = (size_t)decltype(cg)::value << LOGS;
          op[off] = shoup_mul(v.v, __ldg(f2 + off)).v;
        });
      }
    } else {
      lean_round<K, R0, LO0, NT>(w, tw, gload, [&](auto cg, Fp v) { sp[decltype(cg)::value * 32] = v.v; });
    }
  }
  // ---- middle round (three-round levels only) ----
  if constexpr (NR == 3) {
    __syncthreads();
    constexpr int R1 = round_bits(K, 1), LO1 = round_lo(K, 1);
    u32* __restrict__ sp = tile + lean_ins<R1, LO1>(w) * 32 + lane;
    lean_round<K, R1, LO1, NT>(w, tw, [&](auto cg) { return fp_raw(sp[decltype(cg)::value * 32]); },
                               [&](auto cg, Fp v) { sp[decltype(cg)::value * 32] = v.v; });
  }
  // ---- last round: to global memory (DIF level) or to the staging tile (final DIT level) ----
  if constexpr (NR >= 2) {
    __syncthreads();
    constexpr int RL = round_bits(K, NR - 1);
    const u32 sw = lean_ins<RL, 0>(w);
    const u32* __restrict__ sp = tile + sw * 32 + lane;
    auto sload = [&](auto cg) { return fp_raw(sp[decltype(cg)::value * 32]); };
    if constexpr (FINAL) {
      u32* __restrict__ st = stage + lane * P + sw;
      lean_round<K, RL, 0, NT>(w, tw, sload, [&](auto cg, Fp v) { st[decltype(cg)::value] = v.v; });
    } else {
      u32* __restrict__ op = out + ((size_t)sw << LOGS) + t0 + lane;
      const uint2* __restrict__ f2 = a.four + ((size_t)sw << LOGS) + t0 + lane;
      lean_round<K, RL, 0, NT>(w, tw, sload, [&](auto cg, Fp v) {
        constexpr size_t off = (size_t)decltype(cg)::value << LOGS;
        op[off] = shoup_mul(v.v, __ldg(f2 + off)).v;
      });
    }
  }
  if constexpr (FINAL) {
    __syncthreads();
    // slot o of offset t0 + q holds natural index bitrev_K(o) * 2^LOGS + t0 + q, which the committed
    // (bit-reversed) order puts at bitrev_LOGS(t0 + q) * 2^K + o: rows of 2^K words, 128 bytes per warp.
    const u32 o = threadIdx.x & ((1u << K) - 1), h = threadIdx.x >> K;                  // q = h + 2k
    const u32 brt0 = __brev(t0) >> (32 - LOGS);                                         // t0 is a multiple of 32
    const u32* __restrict__ sp = stage + h * P + o;
    u32* __restrict__ op = out + (((size_t)brt0 + ((size_t)h << (LOGS - 1))) << K) + o;
#pragma unroll
    for (int k = 0; k < 16; k++) {
      const u32 br4 = ((k & 1) << 3) | ((k & 2) << 1) | ((k & 4) >> 1) | ((k & 8) >> 3);
      op[(size_t)br4 << (LOGS - 5 + K)] = sp[2 * k * P];
    }
  }
}


// ---- lean contiguous level ---------------------------------------------------------------------
// The fused middle launch of the coset LDE (contiguous inverse DIF -> x shift^i / n -> forward DIT per
// coset, see ntt.cu) on a tile of 4096 CONSECUTIVE elements of one column = 2^(12-K) groups of 2^K,
// 256 threads x 16 elements, with all sizes compile-time constants.  Element e of the tile (group in the
// high bits) lives at e + (e >> 5) in shared memory (one pad word per 32).  In every radix round a
// thread owns the 12 - R bits of e outside the butterfly bits [LO, LO + R): its 5 LANE bits go to the
// five lowest free bit positions that are distinct mod 5 - their address weights are then {1,2,4,8,16}
// mod 32, i.e. every warp access is bank-conflict free - the warp bits and the unrolled group index
// fill the rest.  All addresses are therefore one per-thread register plus an immediate.
struct LeanContigArgs {
  const u32* in; u32* out;
  size_t in_stride, out_stride;          // elements between columns
  size_t out_coset_stride;
  const uint2* small_tw;
  const uint2* scale;                    // per coset: n Shoup pairs, shift_c^bitrev(p) / n at position p
  int ncoset;
};

// destination bit (in the 12-bit tile index) of owner bit `o`: owners 0-4 lane, 5-7 warp, 8.. group counter
template <int R, int LO>
__host__ __device__ constexpr int cmap_dest(int o) {
  bool used[12] = {};
  for (int i = LO; i < LO + R; i++) used[i] = true;
  int dest[12] = {};
  for (int c = 0; c < 5; c++) {                       // lane bits: lowest free bit of every class mod 5
    for (int i = c; i < 10; i += 5)
      if (!used[i]) { dest[c] = i; used[i] = true; break; }
  }
  // sort the five lane destinations so that lane bit 0 gets the lowest (coalescing along the lowest bits)
  for (int a = 0; a < 5; a++)
    for (int b = a + 1; b < 5; b++)
      if (dest[b] < dest[a]) { int t = dest[a]; dest[a] = dest[b]; dest[b] = t; }
  int n = 5;
  for (int i = 0; i < 12; i++)
    if (!used[i]) { dest[n++] = i; used[i] = true; }
  return dest[o];
}
template <int R, int LO, int B>
__host__ __device__ constexpr int cmap_runlen() {
  int len = 1;
  while (B + len < 8 && cmap_dest<R, LO>(B + len) == cmap_dest<R, LO>(B) + len) len++;
  return len;
}
template <int R, int LO, int B>
__device__ __forceinline__ u32 cmap_thread_bits(u32 tid) {
  if constexpr (B >= 8) return 0u;
  else {
    constexpr int d = cmap_dest<R, LO>(B);
    constexpr int len = cmap_runlen<R, LO, B>();
    return (((tid >> B) & ((1u << len) - 1u)) << d) | cmap_thread_bits<R, LO, B + len>(tid);
  }
}
template <int R, int LO>
__host__ __device__ constexpr u32 cmap_group_bits(u32 g) {
  u32 e = 0;
  for (int j = 0; j < 4 - R; j++) e |= ((g >> j) & 1u) << cmap_dest<R, LO>(8 + j);
  return e;
}
__host__ __device__ constexpr u32 cpad(u32 e) { return e + (e >> 5); }
__host__ __device__ constexpr u32 twpad(u32 i) { return i + (i >> 4); }

template <int K, int R, int LO, bool DIF, u32 BLC>
__device__ __forceinline__ void contig_butterflies(Fp* x, u32 blw, const uint2* __restrict__ tw) {
#pragma unroll
  for (int j = 0; j < R; j++) {
    const int lh = DIF ? (R - 1 - j) : j;
    const int sh = K - 1 - LO - lh;
    const uint2* __restrict__ twl = tw + twpad(blw << sh);
#pragma unroll
    for (int p = 0; p < (1 << (R - 1)); p++) {
      const int k = p & ((1 << lh) - 1);
      const int m0 = ((p >> lh) << (lh + 1)) | k;
      const int m1 = m0 + (1 << lh);
      const uint2 w = twl[twpad(BLC << sh) + twpad((u32)k << (LO + sh))];
      if (DIF) {
        const Fp a = x[m0], b = x[m1];
        x[m0] = a + b;
        x[m1] = shoup_mul(a.v - b.v + KB_P, w);
      } else {
        const Fp t = shoup_mul(x[m1].v, w);
        x[m1] = x[m0] - t;
        x[m0] = x[m0] + t;
      }
    }
  }
}

// one radix round; load(ec) / store(ec, v): ec = compile-time part of the tile index, ew = thread part
template <int K, int R, int LO, bool DIF, class Load, class Store>
__device__ __forceinline__ void contig_round(u32 ew, const uint2* __restrict__ tw, Load load, Store store) {
  constexpr int G = 16 >> R;
  constexpr u32 slot_mask = (1u << K) - 1u, lo_mask = (1u << LO) - 1u;
  const u32 blw = ew & slot_mask & lo_mask;
  Fp x[G][1 << R];
  lean_for<0, G>([&](auto gc) {
    constexpr u32 eg = cmap_group_bits<R, LO>(decltype(gc)::value);
    lean_for<0, (1 << R)>([&](auto mc) {
      x[decltype(gc)::value][decltype(mc)::value] = load(std::integral_constant<u32, (eg | ((u32) decltype(mc)::value << LO))>());
    });
  });
  lean_for<0, G>([&](auto gc) {
    constexpr u32 eg = cmap_group_bits<R, LO>(decltype(gc)::value);
    contig_butterflies<K, R, LO, DIF, (eg & slot_mask & lo_mask)>(x[decltype(gc)::value], blw, tw);
    lean_for<0, (1 << R)>([&](auto mc) {
      store(std::integral_constant<u32, (eg | ((u32) decltype(mc)::value << LO))>(), x[decltype(gc)::value][decltype(mc)::value]);
    });
  });
}

template <int K>
__host__ __device__ constexpr size_t lean_contig_smem() {
  return ((size_t)2 * (twpad(1u << (K - 1)) + 2) * 2 + (size_t)2 * (4096 + 128)) * sizeof(u32);
}

template <int K>
__global__ void __launch_bounds__(256, 4) ntt_contig_lean_kernel(LeanContigArgs a, int logn) {
  static_assert(K >= 8 && K <= 12, "lean contiguous level: 4096-element tiles, two or three radix rounds");
  constexpr int NR = (K + 3) / 4;
  constexpr u32 TWW = (twpad(1u << (K - 1)) + 2) * 2;           // words per padded twiddle table
  extern __shared__ __align__(16) u32 smem[];
  uint2* tw_inv = reinterpret_cast<uint2*>(smem);
  uint2* tw_fwd = reinterpret_cast<uint2*>(smem + TWW);
  u32* bufA = smem + 2 * TWW;
  u32* bufB = bufA + 4096 + 128;
  const u32 tid = threadIdx.x;
  const size_t tile0 = (size_t)blockIdx.x << 12;                 // first element of the tile in its column
  const u32* __restrict__ gin = a.in + (size_t)blockIdx.y * a.in_stride + tile0;
  {
    constexpr u32 half = 1u << (K - 1);
    const uint2* src_f = a.small_tw + (half - 1);
    const uint2* src_i = a.small_tw + ((1u << 12) - 1) + (half - 1);
    for (u32 m = tid; m < half; m += 256) { tw_inv[twpad(m)] = src_i[m]; tw_fwd[twpad(m)] = src_f[m]; }
  }
  __syncthreads();

  // ---- inverse DIF, top bits first; the coefficients (bit-reversed) stay in bufA ----
  constexpr int R0 = round_bits(K, 0), LO0 = round_lo(K, 0);
  constexpr int RL = round_bits(K, NR - 1);
  {
    const u32 ew = cmap_thread_bits<R0, LO0, 0>(tid);
    const u32* __restrict__ gp = gin + ew;
    u32* __restrict__ sp = bufA + cpad(ew);
    contig_round<K, R0, LO0, true>(ew, tw_inv, [&](auto ec) { return fp_raw(gp[decltype(ec)::value]); },
                                   [&](auto ec, Fp v) { sp[cpad(decltype(ec)::value)] = v.v; });
  }
  __syncthreads();
  if constexpr (NR == 3) {
    constexpr int R1 = round_bits(K, 1), LO1 = round_lo(K, 1);
    const u32 ew = cmap_thread_bits<R1, LO1, 0>(tid);
    u32* __restrict__ sp = bufA + cpad(ew);
    contig_round<K, R1, LO1, true>(ew, tw_inv, [&](auto ec) { return fp_raw(sp[cpad(decltype(ec)::value)]); },
                                   [&](auto ec, Fp v) { sp[cpad(decltype(ec)::value)] = v.v; });
    __syncthreads();
  }
  {
    const u32 ew = cmap_thread_bits<RL, 0, 0>(tid);
    u32* __restrict__ sp = bufA + cpad(ew);
    contig_round<K, RL, 0, true>(ew, tw_inv, [&](auto ec) { return fp_raw(sp[cpad(decltype(ec)::value)]); },
                                 [&](auto ec, Fp v) { sp[cpad(decltype(ec)::value)] = v.v; });
  }
  __syncthreads();

  // ---- per coset: scale the coefficients, forward DIT bottom bits first, store ----
  for (int c = 0; c < a.ncoset; c++) {
    const uint2* __restrict__ sc = a.scale + ((size_t)c << logn) + tile0;
    u32* __restrict__ gout = a.out + (size_t)c * a.out_coset_stride + (size_t)blockIdx.y * a.out_stride + tile0;
    {
      const u32 ew = cmap_thread_bits<RL, 0, 0>(tid);
      const u32* __restrict__ sp = bufA + cpad(ew);
      const uint2* __restrict__ scp = sc + ew;
      u32* __restrict__ dp = bufB + cpad(ew);
      contig_round<K, RL, 0, false>(ew, tw_fwd,
                                    [&](auto ec) { return shoup_mul(sp[cpad(decltype(ec)::value)], __ldg(scp + decltype(ec)::value)); },
                                    [&](auto ec, Fp v) { dp[cpad(decltype(ec)::value)] = v.v; });
    }
    __syncthreads();
    if constexpr (NR == 3) {
      constexpr int R1 = round_bits(K, 1), LO1 = round_lo(K, 1);
      const u32 ew = cmap_thread_bits<R1, LO1, 0>(tid);
      u32* __restrict__ sp = bufB + cpad(ew);
      contig_round<K, R1, LO1, false>(ew, tw_fwd, [&](auto ec) { return fp_raw(sp[cpad(decltype(ec)::value)]); },
                                      [&](auto ec, Fp v) { sp[cpad(decltype(ec)::value)] = v.v; });
      __syncthreads();
    }
    {
      const u32 ew = cmap_thread_bits<R0, LO0, 0>(tid);
      const u32* __restrict__ sp = bufB + cpad(ew);
      u32* __restrict__ gp = gout + ew;
      contig_round<K, R0, LO0, false>(ew, tw_fwd, [&](auto ec) { return fp_raw(sp[cpad(decltype(ec)::value)]); },
                                      [&](auto ec, Fp v) { gp[decltype(ec)::value] = v.v; });
    }
    __syncthreads();          // bufB is rewritten by the next coset
  }
}

}  // namespace zkb

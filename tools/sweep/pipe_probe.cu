// Issue-rate probe for the integer instructions the prover's kernels are made of (sm_100a):
// per op, 8 independent dependency chains per thread, 16 warps per scheduler, result in
// thread-instructions per clock per SM.  Mixed kernels interleave two ops 1:1 to show whether they
// share a pipe.  One JSON line per probe.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32;
typedef uint64_t u64;

#define CHAINS 8
#define ITERS 16384

template <int OP>
__device__ __forceinline__ void step(u32& x, u64& w, u32 k1, u32 k2) {
  if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(k1), "r"(k2));                       // IMAD
  if (OP == 1) asm volatile("{.reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(w) : "r"(k1));   // IMAD.WIDE
  if (OP == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(k1), "r"(k2));                       // IMAD.HI
  if (OP == 3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x) : "r"(k1), "r"(k2));   // IADD3
  if (OP == 4) asm volatile("{.reg .u32 t; add.u32 t, %0, 0x80ffffff; min.u32 %0, t, %0;}" : "+r"(x));         // VIADDMNMX
  if (OP == 5) asm volatile("min.u32 %0, %0, %1;" : "+r"(x) : "r"(k1));                                        // VIMNMX
  if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(k1), "r"(k2));                    // LOP3
  if (OP == 7) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(k1), "r"(k2));                    // SHF
  if (OP == 8) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(k1));                                        // 2-input add (ptxas picks)
  if (OP == 10) { u64 t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x), "r"(k1)); x = (u32)t ^ (u32)(t >> 32); }     // IMAD.WIDE (RZ addend) + LOP3
  if (OP == 11) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(k1));                                    // IMAD.HI (RZ addend)
  if (OP == 12) {   // Montgomery product x = x * k1 / 2^32 mod p, plus form (kb31.cuh)
    u64 t; u32 m;
    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x), "r"(k1));
    asm volatile("mul.lo.u32 %0, %1, 0x7effffff;" : "=r"(m) : "r"((u32)t));
    u32 r = (u32)((t + (u64)m * 0x7f000001u) >> 32);
    u32 u = r - 0x7f000001u; x = u < r ? u : r;
  }
  if (OP == 13) {   // Shoup product by the constant pair (k1, k2)
    u32 q; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(x), "r"(k2));
    u32 r = x * k1 - q * 0x7f000001u;
    u32 u = r - 0x7f000001u; x = u < r ? u : r;
  }
  if (OP == 14) {   // modular add of a constant
    u32 r = x + k1; u32 u = r - 0x7f000001u; x = u < r ? u : r;
  }
  if (OP == 15) {   // cube: lazy square then product (poseidon2.cuh p2_cube)
    u64 t; u32 m;
    asm volatile("mul.wide.u32 %0, %1, %1;" : "=l"(t) : "r"(x));
    asm volatile("mul.lo.u32 %0, %1, 0x7effffff;" : "=r"(m) : "r"((u32)t));
    u32 y = (u32)((t + (u64)m * 0x7f000001u) >> 32);
    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(y), "r"(x));
    asm volatile("mul.lo.u32 %0, %1, 0x7effffff;" : "=r"(m) : "r"((u32)t));
    u32 r = (u32)((t + (u64)m * 0x7f000001u) >> 32);
    u32 u = r - 0x7f000001u; x = u < r ? u : r;
  }
  if (OP == 9) asm volatile("{.reg .pred q; .reg .u32 t; sub.u32 t, %0, %1; setp.lt.u32 q, %0, %1; selp.u32 %0, %0, t, q;}" : "+r"(x) : "r"(k1));  // sub+setp+sel
}

template <int OPA, int OPB>
__global__ void __launch_bounds__(512, 4) probe(u32* out, u32 k1, u32 k2, long long* clk) {
  u32 x[CHAINS];
  u64 w[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; c++) { x[c] = threadIdx.x * 7 + c; w[c] = x[c]; }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CHAINS; c++) {
      step<OPA>(x[c], w[c], k1, k2);
      if (OPB >= 0) step<OPB < 0 ? 0 : OPB>(x[(c + 4) % CHAINS], w[(c + 4) % CHAINS], k1, k2);
    }
  }
  long long t1 = clock64();
  u32 acc = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; c++) acc += x[c] + (u32)w[c] + (u32)(w[c] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

static const char* NAMES[] = {"IMAD", "IMAD.WIDE", "IMAD.HI", "IADD3", "VIADDMNMX", "VIMNMX", "LOP3", "SHF", "ADD2", "SUB+SETP+SEL",
                              "MULWIDE+LOP", "MULHI", "MONTMUL", "SHOUPMUL", "MODADD", "CUBE"};

template <int OPA, int OPB>
void run(u32* out, long long* clk, int sms) {
  const int blocks = sms * 4, threads = 512;   // 2048 threads per SM = 16 warps per scheduler
  for (int i = 0; i < 3; i++) probe<OPA, OPB><<<blocks, threads>>>(out, 3, 5, clk);
  cudaDeviceSynchronize();
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, probe<OPA, OPB>, threads, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  probe<OPA, OPB><<<blocks, threads>>>(out, 3, 5, clk);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h[1024];
  cudaMemcpy(h, clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < blocks; i++) avg += (double)h[i];
  avg /= blocks;
  const double per_thread = (double)ITERS * CHAINS * (OPB >= 0 ? 2 : 1);
  // 4 CTAs of 512 threads share an SM for the whole run: thread-instructions per clock per SM
  const double per_clk_sm = per_thread * 2048.0 / avg;
  const double by_events = per_thread * blocks * threads / (ms * 1e-3) / sms / 1.965e9;   // assumes 1965 MHz
  printf("{\"probe\": \"pipe\", \"a\": \"%s\", \"b\": \"%s\", \"ms\": %.4f, \"cycles\": %.0f, \"ctas_per_sm\": %d, \"thread_instr_per_clk_per_sm\": %.2f, \"by_events_at_1965MHz\": %.2f}\n",
         NAMES[OPA], OPB >= 0 ? NAMES[OPB < 0 ? 0 : OPB] : "-", ms, avg, occ, per_clk_sm, by_events);
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  u32* out; long long* clk;
  cudaMalloc(&out, (size_t)sms * 4 * 512 * 4);
  cudaMalloc(&clk, sizeof(long long) * 1024);
  run<0, -1>(out, clk, sms); run<1, -1>(out, clk, sms); run<2, -1>(out, clk, sms); run<3, -1>(out, clk, sms);
  run<4, -1>(out, clk, sms); run<5, -1>(out, clk, sms); run<6, -1>(out, clk, sms); run<7, -1>(out, clk, sms);
  run<8, -1>(out, clk, sms); run<9, -1>(out, clk, sms);
  run<10, -1>(out, clk, sms); run<11, -1>(out, clk, sms); run<12, -1>(out, clk, sms); run<13, -1>(out, clk, sms);
  run<14, -1>(out, clk, sms); run<15, -1>(out, clk, sms);
  run<12, 14>(out, clk, sms);  // one product + one modular add
  run<15, 14>(out, clk, sms);  // one cube + one modular add
  run<13, 14>(out, clk, sms);
  run<0, 3>(out, clk, sms);   // IMAD + IADD3: different pipes -> should add up
  run<0, 4>(out, clk, sms);   // IMAD + VIADDMNMX
  run<1, 3>(out, clk, sms);   // IMAD.WIDE + IADD3
  run<1, 4>(out, clk, sms);   // IMAD.WIDE + VIADDMNMX
  run<2, 4>(out, clk, sms);   // IMAD.HI + VIADDMNMX
  run<3, 4>(out, clk, sms);   // IADD3 + VIADDMNMX: same pipe?
  run<0, 1>(out, clk, sms);   // IMAD + IMAD.WIDE
  run<3, 6>(out, clk, sms);   // IADD3 + LOP3
  run<4, 6>(out, clk, sms);   // VIADDMNMX + LOP3
  printf("{\"probe\": \"pipe\", \"cuda\": \"%s\"}\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}

// Stand-alone throughput probe for the Poseidon2 permutation: the same header the library uses,
// compiled once per (ZKB_P2_NH, ZKB_P2_INT_HEAVY) pipe assignment (tools/sweep/Makefile), so one
// GPU call ranks the variants.  Prints one JSON line: G permutations/s and a host cross-check.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "poseidon2.cuh"

namespace zkb {
const P2Consts& p2_host_consts() {
  static const P2Consts c = p2_make_consts();
  return c;
}
__constant__ P2Consts d_p2;
}  // namespace zkb
using namespace zkb;

__global__ void __launch_bounds__(128, 16) chain_kernel(u32* states, size_t n, int iters) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fp st[16];
#pragma unroll
  for (int j = 0; j < 16; j++) st[j] = fp_raw(states[j * n + i]);
  for (int it = 0; it < iters; it++) p2_permute_with(st, d_p2);
#pragma unroll
  for (int j = 0; j < 16; j++) states[j * n + i] = st[j].v;
}

int main(int argc, char** argv) {
  const size_t n = (size_t)1 << (argc > 1 ? atoi(argv[1]) : 21);
  const int iters = argc > 2 ? atoi(argv[2]) : 8;
  cudaMemcpyToSymbol(d_p2, &p2_host_consts(), sizeof(P2Consts));
  std::vector<u32> h(16 * n);
  u32 x = 12345;
  for (auto& v : h) { x = x * 1664525u + 1013904223u; v = (x >> 1) % KB_P; }
  // edge values in the first states
  for (int j = 0; j < 16; j++) { h[j * n + 0] = 0; h[j * n + 1] = KB_P - 1; h[j * n + 2] = (j & 1) ? KB_P - 1 : 0; h[j * n + 3] = 1; }
  u32* d;
  cudaMalloc(&d, 16 * n * 4);
  cudaMemcpy(d, h.data(), 16 * n * 4, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const unsigned blocks = (unsigned)((n + 127) / 128);
  chain_kernel<<<blocks, 128>>>(d, n, iters);   // warm-up (also the checked run)
  std::vector<u32> out(16 * n);
  cudaMemcpy(out.data(), d, 16 * n * 4, cudaMemcpyDeviceToHost);
  size_t bad = 0;
  for (size_t i = 0; i < 4096 && i < n; i++) {
    Fp st[16];
    for (int j = 0; j < 16; j++) st[j] = fp_raw(h[j * n + i]);
    for (int it = 0; it < iters; it++) p2_permute_host(st);
    for (int j = 0; j < 16; j++) bad += st[j].v != out[j * n + i];
  }
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    chain_kernel<<<blocks, 128>>>(d, n, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaError_t err = cudaDeviceSynchronize();
  printf("{\"probe\": \"p2\", \"nh\": %d, \"int_heavy\": %d, \"n\": %zu, \"iters\": %d, \"ms\": %.4f, \"gperm_s\": %.4f, \"mismatch\": %zu, \"cuda\": \"%s\"}\n",
         ZKB_P2_NH, ZKB_P2_INT_HEAVY, n, iters, best, (double)n * iters / best * 1e-6, bad, cudaGetErrorString(err));
  return bad != 0;
}

// A CPU stand-in for the slice of the CUDA runtime and execution model that csrc/tracegen.cu, csrc/derive.cu and
// csrc/machine.cpp use, so that the LAUNCHERS and __global__ wrappers of the row fillers (K6, K6b, K6c) and of K7 - grid
// sizes, shared-memory tiles, __syncthreads phases, bounds checks, scratch buffers, the recursion of the curve-point scan -
// run in the CPU suite from the same source text (tests/cudaemu/build.py rewrites only the `k<<<g, b, 0, s>>>(args)` syntax
// into emu_launch(k, g, b, args)).  Test infrastructure only; the product never sees this header (it is found before the
// real <cuda_runtime.h> only through -I tests/cudaemu).
//
// Execution model: the blocks of a grid run one after the other; the threads of a block are real threads that pass a baton,
// so exactly one runs at a time and the others wait - at the start, or inside __syncthreads(), where a thread hands the baton
// on and waits at the block's barrier.  `__shared__` is `static` (one block at a time), `__constant__` a plain global.
#pragma once
#include <barrier>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
typedef struct EmuStream* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
constexpr unsigned cudaHostAllocMapped = 2;

inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
// device memory is not zero either: fresh allocations carry a pattern
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); if (*p) memset(*p, 0xA5, n); return *p ? 0 : 2; }
inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
inline cudaError_t cudaFree(void* p) { free(p); return 0; }
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { free(p); return 0; }
inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
#define cudaMemcpyToSymbol(sym, src, n) (memcpy((void*)&(sym), (src), (n)), cudaSuccess)

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__
#define __shared__ static
#define __launch_bounds__(...)

struct EmuIdx { unsigned x = 0, y = 0, z = 0; };
struct EmuBlock {
  std::vector<void*> poisoned;      // shared arrays already filled with the poison pattern for this block
  std::barrier<> barrier;
  explicit EmuBlock(unsigned n) : barrier((std::ptrdiff_t)n) {}
};
inline thread_local EmuIdx threadIdx, blockIdx, blockDim, gridDim;
inline thread_local EmuBlock* emu_block = nullptr;
inline std::mutex emu_baton;

inline void __syncthreads() {
  emu_baton.unlock();
  emu_block->barrier.arrive_and_wait();
  emu_baton.lock();
}

// tests/cudaemu/build.py puts this after every `__shared__` array declaration (runs under the baton)
inline void emu_poison(void* p, size_t bytes) {
  for (void* q : emu_block->poisoned) if (q == p) return;
  emu_block->poisoned.push_back(p);
  unsigned* w = static_cast<unsigned*>(p);
  for (size_t i = 0; i < bytes / 4; i++) w[i] = 0xDEADBEEFu;
}

template <class K, class... A>
void emu_launch(K kernel, unsigned grid, unsigned block, A... args) {
  // `block` threads per launch, reused for every block of the grid; a block starts when the one before it has ended
  std::vector<std::unique_ptr<EmuBlock>> blocks;
  for (unsigned b = 0; b < grid; b++) blocks.emplace_back(new EmuBlock(block));
  std::barrier<> block_end((std::ptrdiff_t)block);
  std::vector<std::thread> ts;
  ts.reserve(block);
  for (unsigned t = 0; t < block; t++)
    ts.emplace_back([&, t] {
      for (unsigned b = 0; b < grid; b++) {
        threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
        emu_block = blocks[b].get();
        emu_baton.lock();
        kernel(args...);
        emu_baton.unlock();
        emu_block->barrier.arrive_and_drop();
        block_end.arrive_and_wait();
      }
    });
  for (auto& th : ts) th.join();
}

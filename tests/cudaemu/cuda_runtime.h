// A CPU stand-in for the slice of the CUDA runtime and execution model that csrc/tracegen.cu, csrc/derive.cu and
// csrc/machine.cpp use, so that the LAUNCHERS and __global__ wrappers of the row fillers (K6, K6b, K6c) and of K7 - grid
// sizes, shared-memory tiles, __syncthreads phases, bounds checks, scratch buffers, the recursion of the curve-point scan -
// run in the CPU suite from the same source text (tests/cudaemu/build.py rewrites only the `k<<<g, b, 0, s>>>(args)` syntax
// into emu_launch(k, g, b, args)).  Test infrastructure only; the product never sees this header (it is found before the
// real <cuda_runtime.h> only through -I tests/cudaemu).
//
// Execution model: the blocks of a grid run one after the other; the threads of a block are fibers of the launching thread:
// one runs at a time, until it ends or reaches a barrier (__syncthreads, a warp shuffle), and a barrier opens when every live
// thread of the block / warp has arrived.  `__shared__` is `static` (one block at a time), `__constant__` a plain global.
#pragma once
#include <ucontext.h>
#include <algorithm>
#include <array>
#include <cstdio>
#include <functional>
#include <cstdint>
#include <map>
#include <type_traits>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0, cudaErrorEmulated = 999;
typedef struct EmuStream* cudaStream_t;
typedef struct EmuEvent* cudaEvent_t;
typedef struct EmuPool* cudaMemPool_t;
typedef struct EmuLibrary* cudaLibrary_t;
typedef struct EmuKernel* cudaKernel_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaMemPoolAttr { cudaMemPoolReuseAllowInternalDependencies = 3, cudaMemPoolAttrReleaseThreshold = 4 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0, cudaDriverEntryPointSymbolNotFound = 1 };
constexpr unsigned cudaHostAllocMapped = 2, cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEnableDefault = 0;

struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

// ---- memory: everything is host memory; a registry remembers what kind each allocation stands for -------------------------
struct EmuAllocs {
  std::mutex mu;
  std::map<char*, std::pair<size_t, cudaMemoryType>> m;
  void add(void* p, size_t n, cudaMemoryType t) { std::lock_guard<std::mutex> l(mu); m[(char*)p] = {n ? n : 1, t}; }
  void remove(void* p) { std::lock_guard<std::mutex> l(mu); m.erase((char*)p); }
  cudaMemoryType type_of(const void* p) {
    std::lock_guard<std::mutex> l(mu);
    auto it = m.upper_bound((char*)p);
    if (it == m.begin()) return cudaMemoryTypeUnregistered;
    --it;
    return (char*)p < it->first + it->second.first ? it->second.second : cudaMemoryTypeUnregistered;
  }
};
inline EmuAllocs emu_allocs;
inline cudaError_t emu_alloc(void** p, size_t n, cudaMemoryType t) {
  *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256);
  if (!*p) return 2;
  memset(*p, 0xA5, n);          // device memory is not zero: fresh allocations carry a pattern
  emu_allocs.add(*p, n, t);
  return cudaSuccess;
}
inline cudaError_t emu_free(void* p) { if (p) { emu_allocs.remove(p); free(p); } return cudaSuccess; }

inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { return emu_alloc(p, n, cudaMemoryTypeDevice); }
inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return emu_alloc(p, n, cudaMemoryTypeDevice); }
inline cudaError_t cudaFree(void* p) { return emu_free(p); }
inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return emu_free(p); }
inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return emu_alloc(p, n, cudaMemoryTypeHost); }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return emu_alloc(p, n, cudaMemoryTypeHost); }
inline cudaError_t cudaFreeHost(void* p) { return emu_free(p); }
inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) {
  if (emu_allocs.type_of(h) != cudaMemoryTypeHost) return cudaErrorEmulated;
  *d = h;
  return cudaSuccess;
}
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
  a->type = emu_allocs.type_of(p); a->device = 0; a->devicePointer = (void*)p; a->hostPointer = (void*)p;
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                                     cudaStream_t) {
  for (size_t r = 0; r < height; r++) memmove((char*)d + r * dpitch, (const char*)s + r * spitch, width);
  return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = 0; *t = 0; return cudaSuccess; }     // nothing to reserve
#define cudaMemcpyToSymbol(sym, src, n) (memcpy((void*)&(sym), (src), (n)), cudaSuccess)

// ---- devices, streams, events: every call completes before it returns, so ordering primitives have nothing to do -------------
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }      // four "SMs"
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = nullptr; return cudaSuccess; }
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
template <class F> cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }
// no driver: stream memory operations (the pull mode) and run-time compiled kernels are reported unavailable, the library
// falls back to DMA uploads and to the data-driven K3 / K5 kernels
inline cudaError_t cudaGetDriverEntryPoint(const char*, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* q = nullptr) {
  *fn = nullptr;
  if (q) *q = cudaDriverEntryPointSymbolNotFound;
  return cudaErrorEmulated;
}
inline cudaError_t cudaLibraryLoadData(cudaLibrary_t*, const void*, void*, void*, unsigned, void*, void*, unsigned) { return cudaErrorEmulated; }
inline cudaError_t cudaLibraryGetKernel(cudaKernel_t*, cudaLibrary_t, const char*) { return cudaErrorEmulated; }
inline cudaError_t cudaLaunchKernel(const void*, dim3, dim3, void**, size_t, cudaStream_t) { return cudaErrorEmulated; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

// ---- device-side intrinsics ----------------------------------------------------------------------------------------------
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcs(const T* p) { return *p; }
inline unsigned __brev(unsigned x) { unsigned r = 0; for (int i = 0; i < 32; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline void __threadfence() {}
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
template <class A, class B> inline typename std::common_type<A, B>::type min(A a, B b) { using C = typename std::common_type<A, B>::type; return (C)a < (C)b ? (C)a : (C)b; }
template <class A, class B> inline typename std::common_type<A, B>::type max(A a, B b) { using C = typename std::common_type<A, B>::type; return (C)a < (C)b ? (C)b : (C)a; }
// one thread runs at a time, so plain read-modify-writes are atomic here
template <class T, class V> inline T atomicAdd(T* p, V v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class V> inline T atomicMin(T* p, V v) { T o = *p; if ((T)v < o) *p = (T)v; return o; }
template <class T> inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

struct EmuIdx { unsigned x = 0, y = 0, z = 0; };
// One CUDA thread = one fiber (ucontext) of the OS thread that launches the grid.  A fiber runs until it ends or reaches a
// barrier (__syncthreads: the block's; a warp shuffle: its warp's), the scheduler then runs the next one; a barrier opens when
// every fiber that is still alive in its block / warp has arrived.
struct EmuFiber {
  ucontext_t ctx;
  int wait = 0;          // 0 runnable, 1 at the block barrier, 2 at the warp barrier
  bool done = false;
};
struct EmuBlock {
  std::vector<void*> poisoned;      // shared arrays already filled with the poison pattern for this block
  std::vector<EmuFiber> fibers;
  std::vector<std::array<unsigned, 32>> warp_vals;
  ucontext_t scheduler;
  unsigned current = 0;
};
inline thread_local EmuIdx threadIdx, blockIdx, blockDim, gridDim;
inline thread_local EmuBlock* emu_block = nullptr;
inline thread_local void* emu_dyn_smem = nullptr;
inline thread_local std::function<void()>* emu_body = nullptr;
inline std::mutex emu_launch_mu;       // one grid at a time: `__shared__` arrays are statics

inline void emu_yield(int wait) {
  EmuFiber& f = emu_block->fibers[emu_block->current];
  f.wait = wait;
  swapcontext(&f.ctx, &emu_block->scheduler);
}
inline void __syncthreads() { emu_yield(1); }
inline void* emu_dyn_shared() { return emu_dyn_smem; }
// every live lane of the warp takes part (all masks in this library are full)
inline unsigned __shfl_xor_sync(unsigned, unsigned v, int d) {
  const unsigned t = emu_block->current;
  std::array<unsigned, 32>& vals = emu_block->warp_vals[t / 32];
  vals[t % 32] = v;
  emu_yield(2);
  const unsigned r = vals[(t % 32) ^ (unsigned)d];
  emu_yield(2);
  return r;
}
// tests/cudaemu/build.py puts this after every `__shared__` array declaration
inline void emu_poison(void* p, size_t bytes) {
  for (void* q : emu_block->poisoned) if (q == p) return;
  emu_block->poisoned.push_back(p);
  unsigned* w = static_cast<unsigned*>(p);
  for (size_t i = 0; i < bytes / 4; i++) w[i] = 0xDEADBEEFu;
}
inline void emu_trampoline() {
  (*emu_body)();
  EmuFiber& f = emu_block->fibers[emu_block->current];
  f.done = true;
  swapcontext(&f.ctx, &emu_block->scheduler);
}

template <class K, class... A>
void emu_launch_dyn(K kernel, dim3 grid, dim3 block, size_t dyn_bytes, A... args) {
  std::lock_guard<std::mutex> one_grid(emu_launch_mu);
  const unsigned nthreads = block.x * block.y * block.z, nblocks = grid.x * grid.y * grid.z;
  constexpr size_t STACK = 512 << 10;
  std::vector<std::unique_ptr<char[]>> stacks;
  for (unsigned t = 0; t < nthreads; t++) stacks.emplace_back(new char[STACK]);
  std::vector<unsigned> dyn((dyn_bytes + 3) / 4 + 4);
  std::function<void()> body = [&] { kernel(args...); };
  emu_body = &body;
  emu_dyn_smem = (void*)(((uintptr_t)dyn.data() + 15) & ~(uintptr_t)15);
  blockDim.x = block.x; blockDim.y = block.y; blockDim.z = block.z;
  gridDim.x = grid.x; gridDim.y = grid.y; gridDim.z = grid.z;
  EmuBlock blk;
  blk.fibers.resize(nthreads);
  blk.warp_vals.resize((nthreads + 31) / 32);
  emu_block = &blk;
  for (unsigned b = 0; b < nblocks; b++) {
    blockIdx.x = b % grid.x; blockIdx.y = b / grid.x % grid.y; blockIdx.z = b / (grid.x * grid.y);
    blk.poisoned.clear();
    std::fill(dyn.begin(), dyn.end(), 0xDEADBEEFu);
    for (unsigned t = 0; t < nthreads; t++) {
      EmuFiber& f = blk.fibers[t];
      f.wait = 0; f.done = false;
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = stacks[t].get();
      f.ctx.uc_stack.ss_size = STACK;
      f.ctx.uc_link = nullptr;
      makecontext(&f.ctx, emu_trampoline, 0);
    }
    unsigned alive = nthreads;
    while (alive) {
      bool progress = false;
      for (unsigned t = 0; t < nthreads; t++) {
        EmuFiber& f = blk.fibers[t];
        if (f.done || f.wait) continue;
        threadIdx.x = t % block.x; threadIdx.y = t / block.x % block.y; threadIdx.z = t / (block.x * block.y);
        blk.current = t;
        swapcontext(&blk.scheduler, &f.ctx);
        progress = true;
        if (f.done) alive--;
      }
      // barriers open when every live fiber of the block / warp has arrived
      bool all_block = alive > 0;
      for (auto& f : blk.fibers) if (!f.done && f.wait != 1) { all_block = false; break; }
      if (all_block) { for (auto& f : blk.fibers) f.wait = 0; progress = true; }
      for (unsigned w = 0; w * 32 < nthreads; w++) {
        bool all = false, any = false;
        for (unsigned t = 32 * w; t < nthreads && t < 32 * w + 32; t++) {
          const EmuFiber& f = blk.fibers[t];
          if (f.done) continue;
          if (f.wait == 2) any = true; else { any = false; all = false; goto next_warp; }
          all = true;
        }
        if (all && any) { for (unsigned t = 32 * w; t < nthreads && t < 32 * w + 32; t++) if (!blk.fibers[t].done) blk.fibers[t].wait = 0; progress = true; }
      next_warp:;
      }
      if (alive && !progress) { fprintf(stderr, "cudaemu: deadlock (a barrier some thread of the block never reaches)\n"); abort(); }
    }
  }
  emu_block = nullptr;
}
template <class K, class... A>
void emu_launch(K kernel, dim3 grid, dim3 block, A... args) { emu_launch_dyn(kernel, grid, block, 0, args...); }

// Shared host-side plumbing of libzkb200: error handling, stream-ordered device memory,
// a parameter arena for small host->device tables, and the column-major device matrix.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "kb31.cuh"

namespace zkb {

#define ZKB_CUDA(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " +    \
                               __FILE__ + ":" + std::to_string(__LINE__) + " (" #expr ")");        \
  } while (0)

// every kernel launch site goes through this macro: it also feeds the live launch counter that
// bench.py reports as `gpu_launches`
extern std::atomic<unsigned long long> g_kernel_launches;
#define ZKB_CHECK_LAUNCH()                 \
  do {                                     \
    ZKB_CUDA(cudaGetLastError());          \
    g_kernel_launches.fetch_add(1, std::memory_order_relaxed); \
  } while (0)

// Device buffer on the context stream (cudaMallocAsync pool keeps freed blocks cached).
struct DevBuf {
  u32* p = nullptr;
  size_t words = 0;
  cudaStream_t stream = nullptr;
  DevBuf() {}
  DevBuf(size_t n_words, cudaStream_t s) : words(n_words), stream(s) {
    if (n_words) ZKB_CUDA(cudaMallocAsync((void**)&p, n_words * sizeof(u32), s));
  }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), words(o.words), stream(o.stream) { o.p = nullptr; o.words = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; words = o.words; stream = o.stream; o.p = nullptr; o.words = 0; }
    return *this;
  }
  void release() {
    if (p) { cudaFreeAsync(p, stream); p = nullptr; }
    words = 0;
  }
  ~DevBuf() { release(); }
};

// Column-major device matrix: element (r, c) at d[c * height + r].  Columns are the polynomials,
// so NTTs run over contiguous memory and row-wise kernels (leaf hashing, constraint evaluation)
// read with one thread per row fully coalesced.
struct DevMat {
  DevBuf buf;
  size_t height = 0, width = 0;
  u32* d() const { return buf.p; }
  DevMat() {}
  DevMat(size_t h, size_t w, cudaStream_t s) : buf(h * w, s), height(h), width(w) {}
};

// Small host -> device transfers issued on a COMPUTE stream must not go through the copy engine: a
// cudaMemcpyAsync(HostToDevice) queues behind the multi-GB trace upload another shard has in flight on the
// copy stream (one H2D engine, FIFO) and stalls the lane until that upload ends - measured: four shards in
// flight ran as slowly as one (191 ms per shard instead of 123).  So the words are written into pinned,
// device-mapped host memory and a tiny kernel on the lane's stream pulls them over PCIe (layout.cu).
void pull_words(u32* dst_dev, const u32* src_pinned_mapped, size_t n_words, cudaStream_t s);

// Bump arena for small tables handed to kernels (matrix lists, opening schedules...):
// written into pinned host memory, pulled in-stream by a kernel, recycled when the owner knows the stream
// has drained.
struct ParamArena {
  char* host = nullptr;
  char* dev = nullptr;
  size_t cap = 0, used = 0;
  cudaStream_t stream = nullptr;
  void init(size_t bytes, cudaStream_t s) {
    cap = bytes; stream = s;
    ZKB_CUDA(cudaHostAlloc((void**)&host, bytes, cudaHostAllocMapped));
    ZKB_CUDA(cudaMalloc((void**)&dev, bytes));
  }
  void destroy() {
    if (host) cudaFreeHost(host);
    if (dev) cudaFree(dev);
    host = dev = nullptr;
  }
  void reset() { used = 0; }
  template <class T>
  T* push(const T* data, size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
    if (used + bytes > cap) throw std::runtime_error("zkb200: parameter arena exhausted");
    memcpy(host + used, data, n * sizeof(T));
    pull_words((u32*)(dev + used), (const u32*)(host + used), (n * sizeof(T) + 3) / 4, stream);
    T* r = (T*)(dev + used);
    used += bytes;
    return r;
  }
};

static inline unsigned ceil_div(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace zkb

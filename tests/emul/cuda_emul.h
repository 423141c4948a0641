// CPU emulation of the CUDA execution model -- TEST INFRASTRUCTURE ONLY.
//
// The authoring container has nvcc but no GPU.  To debug kernel *logic*
// (indexing, tiling, halos, barriers, shuffles, atomics) without spending GPU
// minutes, tests/emul/build_emul.py compiles the very same .cu sources with g++
// against this header into tests/emul/_build/libgnas_emul.so.  Every CUDA thread is
// a fiber (hand-rolled x86-64 context switch), a block is a set of fibers that
// meet at __syncthreads(), warps exchange registers through a per-warp mailbox,
// and blocks of a launch run on OpenMP threads (cooperative launches: all blocks
// as fibers of one OS thread, so grid barriers can spin).  The product package
// never loads that library: it is reachable only when a test explicitly points
// the ctypes loader at it.  It is not a CPU fallback and is never timed.
#pragma once
#ifndef GNAS_EMUL
#error "cuda_emul.h is only for -DGNAS_EMUL builds"
#endif
#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>
#include <omp.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local
#define __restrict__ __restrict

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_ { unsigned x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(16) int4 { int x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }

typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorLaunchFailure = 719 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
enum { cudaMemcpyDeviceToDevice = 3, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 };

namespace emul {

struct Block;
struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  bool done = false;
  uint3_ tid;
  Block* block = nullptr;
  int lin = 0;  // linear thread id in block
};
struct Warp {
  alignas(16) unsigned char mail[32][16];
  int arrived = 0;
  unsigned gen = 0;
  int nlanes = 32;
};
struct NamedBarrier { int arrived = 0; unsigned gen = 0; };
struct Block {
  uint3_ bid;
  dim3 bdim, gdim;
  int nthreads = 0, alive = 0;
  int arrived = 0;
  unsigned gen = 0;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  unsigned char* dyn_smem = nullptr;
  NamedBarrier named[16];                 // bar.sync id, count
  std::vector<float> tmem;                // tensor memory of the CTA: [128 lanes][512 columns], allocated on demand
};

extern thread_local Fiber* cur;
extern thread_local void* sched_sp;
void yield();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body, bool cooperative);
inline Block* cur_block() { return cur->block; }

inline void syncthreads() {
  Block* b = cur->block;
  unsigned g = b->gen;
  if (++b->arrived >= b->alive) { b->arrived = 0; b->gen++; }
  else while (b->gen == g) yield();
}
inline Warp* my_warp() { return &cur->block->warps[cur->lin >> 5]; }
inline void syncwarp() {
  Warp* w = my_warp();
  unsigned g = w->gen;
  if (++w->arrived >= w->nlanes) { w->arrived = 0; w->gen++; }
  else while (w->gen == g) yield();
}
template <typename T>
inline T shfl_generic(T v, int src_lane) {
  static_assert(sizeof(T) <= 16, "shuffle payload too large");
  Warp* w = my_warp();
  int lane = cur->lin & 31;
  memcpy(w->mail[lane], &v, sizeof(T));
  syncwarp();
  T r = v;
  if (src_lane >= 0 && src_lane < w->nlanes) memcpy(&r, w->mail[src_lane], sizeof(T));
  syncwarp();
  return r;
}
}  // namespace emul

#define threadIdx (emul::cur->tid)
#define blockIdx (emul::cur->block->bid)
#define blockDim (emul::cur->block->bdim)
#define gridDim (emul::cur->block->gdim)
#define warpSize 32

static inline void __syncthreads() { emul::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emul::syncwarp(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_block() {}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
  int lane = emul::cur->lin & 31; (void)width; return emul::shfl_generic(v, lane ^ m);
}
template <typename T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  int lane = emul::cur->lin & 31;
  int src = lane + (int)d;
  if ((lane % width) + (int)d >= width) src = lane;
  return emul::shfl_generic(v, src);
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  int lane = emul::cur->lin & 31;
  return emul::shfl_generic(v, (lane / width) * width + (src % width));
}
template <typename T> static inline T __ldg(const T* p) { return *p; }

template <typename T> static inline T emul_atomic_add(T* addr, T v) {
  T old, neu;
  __atomic_load(addr, &old, __ATOMIC_RELAXED);
  do { neu = old + v; } while (!__atomic_compare_exchange(addr, &old, &neu, true, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
  return old;
}
static inline float atomicAdd(float* a, float v) { return emul_atomic_add(a, v); }
static inline double atomicAdd(double* a, double v) { return emul_atomic_add(a, v); }
static inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicExch(unsigned* a, unsigned v) { return __atomic_exchange_n(a, v, __ATOMIC_SEQ_CST); }
static inline int atomicMax(int* a, int v) {
  int old = __atomic_load_n(a, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(a, &old, v, true, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {}
  return old;
}
static inline unsigned emul_ld_acquire(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }

#define __expf(x) expf(x)
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
using std::max;
using std::min;

#define GNAS_EMUL_YIELD() emul::yield()
#include "tc_emul.h"

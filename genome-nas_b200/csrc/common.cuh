// Shared device/host helpers for the genomeDARTS sm_100a kernels.
#pragma once
#ifdef GNAS_EMUL
// tests/emul/cuda_emul.h is force-included by the emulator build (test infrastructure only)
namespace gnas {
void prof_before(const char* kernel, const char* where, cudaStream_t stream);
void prof_after(cudaStream_t stream);
}
#define GNAS_LAUNCH(kernel, grid, block, smem, stream, ...)                   \
  do {                                                                        \
    gnas::prof_before(#kernel, __PRETTY_FUNCTION__, stream);                  \
    emul::launch(grid, block, smem, [&]() { kernel(__VA_ARGS__); }, false);   \
  } while (0)
#define GNAS_LAUNCH_COOP(kernel, grid, block, smem, stream, ...)              \
  do {                                                                        \
    gnas::prof_before(#kernel, __PRETTY_FUNCTION__, stream);                  \
    emul::launch(grid, block, smem, [&]() { kernel(__VA_ARGS__); }, true);    \
  } while (0)
#define GNAS_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emul::cur_block()->dyn_smem)
#define GNAS_SET_SMEM(kernel, bytes) 0
#define GNAS_SPIN_PAUSE() GNAS_EMUL_YIELD()
#else
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
namespace gnas {
void prof_before(const char* kernel, const char* where, cudaStream_t stream);
void prof_after(cudaStream_t stream);
}
#define GNAS_LAUNCH(kernel, grid, block, smem, stream, ...)          \
  do {                                                               \
    gnas::prof_before(#kernel, __PRETTY_FUNCTION__, stream);         \
    kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);              \
    gnas::prof_after(stream);                                        \
  } while (0)
#define GNAS_DYN_SMEM(T, name)                                       \
  extern __shared__ __align__(16) unsigned char gnas_dyn_smem_raw[]; \
  T* name = reinterpret_cast<T*>(gnas_dyn_smem_raw)
#define GNAS_SET_SMEM(kernel, bytes) \
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#define GNAS_SPIN_PAUSE() __nanosleep(32)
#endif

#define GNAS_OK 0
#define GNAS_ERR_ARG (-1)
#define GNAS_ERR_SHAPE (-2)
#define GNAS_ERR_UNSUPPORTED (-3)
#define GNAS_ERR_WORKSPACE (-4)

namespace gnas {

constexpr int kThreads = 256;
constexpr float kBnEps = 1e-5f;

// how a tensor is transformed while it is read
enum : int { T_NONE = 0, T_RELU = 1, T_BN_RELU = 2, T_BN = 3 };
// where the result of a backward data kernel goes
enum : int { O_STORE = 0, O_RELU_ACC = 1, O_BNRELU_STORE = 2, O_ACC = 3 };

void set_error(const char* msg);
int check_launch(const char* what);

__device__ __forceinline__ float apply_tin(float v, int tin, float mean, float rstd) {
  if (tin == T_RELU) return v > 0.f ? v : 0.f;
  if (tin == T_BN_RELU) { v = (v - mean) * rstd; return v > 0.f ? v : 0.f; }
  if (tin == T_BN) return (v - mean) * rstd;
  return v;
}

// 128-bit reduction into global memory (one L2 atomic transaction instead of four; sm_90+)
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
#ifdef GNAS_EMUL
  atomicAdd(addr, a); atomicAdd(addr + 1, b); atomicAdd(addr + 2, c); atomicAdd(addr + 3, d);
#else
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
#endif
}

// 16-byte asynchronous global -> shared copy (L2 only) and its group fences
#ifdef GNAS_EMUL
__device__ __forceinline__ void gnas_cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
__device__ __forceinline__ void gnas_cp_async_commit() {}
__device__ __forceinline__ void gnas_cp_async_wait_all() {}
#else
__device__ __forceinline__ void gnas_cp_async16(void* dst, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void gnas_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void gnas_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::); }
#endif

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#ifndef GNAS_EMUL
// Kernel attributes (dynamic shared-memory opt-in) are per DEVICE: one table entry per ordinal, so that a process
// which drives several GPUs prepares every kernel on each of them (returns true when the attribute must be set).
constexpr int kMaxDevices = 64;
static inline bool need_prepare(int* table, int want) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return true;
  if (want > table[dev]) { table[dev] = want; return true; }
  return false;
}
#endif

// A second stream per device and purpose for work that is off the critical path of a backward pass (fork / join by
// events, which a CUDA-graph capture of the caller's stream records as plain dependencies).  `which`: 0 = RHN
// backward products (GNAS_RHN_SIDE=0 disables), 1..3 = independent chains of a cell pass (GNAS_WG_SIDE=0 disables).  st == nullptr:
// not available (emulator, disabled, stream creation failed) -- callers then stay on their own stream.
#ifdef GNAS_EMUL      // the emulator has one in-order queue: these are never called
typedef void* cudaEvent_t;
static inline int cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
static inline int cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, int) { return 0; }
#endif
struct SideStream { cudaStream_t st; cudaEvent_t fork; cudaEvent_t join[2]; };
SideStream side_stream(int which);

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return cdiv(a, b) * b; }

}  // namespace gnas

// tcgen05 / TMEM / mbarrier / TMA-bulk PTX wrappers shared by the tensor-core kernels (sm_100a only).
// Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor / InstrDescriptor).
#pragma once
#include "common.cuh"

#include <cstdlib>
#ifndef GNAS_TC_DEFAULT
#define GNAS_TC_DEFAULT 1
#endif
namespace gnas {
// the tcgen05 kernels are the default on the GPU; GNAS_TC=0 selects the SIMT kernels (A/B comparisons, debugging)
static inline bool tc_enabled() {
#ifdef GNAS_EMUL
  return false;
#else
  static int v = -1;
  if (v < 0) { const char* s = getenv("GNAS_TC"); v = s ? atoi(s) : GNAS_TC_DEFAULT; }
  return v != 0;
#endif
}
}  // namespace gnas

#ifndef GNAS_EMUL
namespace gnas {
namespace tc {

constexpr int kTcThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t to_tf32(float x) { uint32_t u; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x)); return u; }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), SWIZZLE_NONE
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ uint32_t make_idesc(int N) {
  // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b format TF32 (2) [7,10)/[10,13), K-major A and B,
  // n_dim = N>>3 [17,23), m_dim = 128>>4 [24,29)
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// same with an explicit M (64 or 128): m_dim = M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc_mn(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// bounded wait: a kernel that would otherwise hang traps instead (the GPU must never be left spinning)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    if (ok) return;
    if (clock64() - t0 > 2000000000LL) __trap();
  }
}
// non-blocking probe of the phase with the given parity
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine; completion is counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
  return pred != 0;
}
// generic-proxy accesses (other SMs' st.global observed through an acquire) -> async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, int cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, int cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                 "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace tc
}  // namespace gnas
#else   // GNAS_EMUL: the same names on the functional model of tests/emul/tc_emul.h (test infrastructure)
namespace gnas {
namespace tc {
inline uint32_t smem_u32(const void* p) { return emul::smem_offset(p); }
inline uint32_t to_tf32(float x) { return emul::tf32_rna(x); }
inline uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
inline uint32_t make_idesc_mn(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
inline void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  emul::mma_tf32(tmem_d, adesc, bdesc, idesc, accumulate != 0);
}
inline void mma_commit(uint64_t* bar) { emul::mbar_arrive(bar); }     // MMAs ran synchronously at issue
inline void mbar_init(uint64_t* bar, uint32_t count) { emul::mbar_init(bar, count); }
inline bool mbar_test(uint64_t* bar, uint32_t parity) { return emul::mbar_test(bar, parity); }
inline void mbar_wait(uint64_t* bar, uint32_t parity) { while (!emul::mbar_test(bar, parity)) emul::yield(); }
inline void mbar_arrive(uint64_t* bar) { emul::mbar_arrive(bar); }
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { emul::mbar_arrive(bar, (int64_t)bytes); }
inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  if ((bytes & 15) || (reinterpret_cast<uintptr_t>(dst) & 15) || (reinterpret_cast<uintptr_t>(src) & 15)) {
    fprintf(stderr, "emul: misaligned bulk copy\n"); abort();
  }
  emul::smem_offset(dst);
  memcpy(dst, src, bytes);
  emul::mbar_complete_tx(bar, (int64_t)bytes);
}
inline bool elect_one() { return (emul::cur->lin & 31) == 0; }
inline void fence_async_smem() {}
inline void fence_proxy_async_all() {}
inline void tc_fence_before() {}
inline void tc_fence_after() {}
inline void mbar_fence_init() {}
inline void named_bar_sync(int id, int nthreads) { emul::named_bar_sync(id, nthreads); }
inline void tmem_alloc(uint32_t* slot, int cols) { (void)cols; emul::tmem_base(); if ((emul::cur->lin & 31) == 0) *slot = 0; }
inline void tmem_dealloc(uint32_t, int) {}
inline void tmem_ld32(uint32_t taddr, float (&v)[32]) { emul::tmem_ld32(taddr, v); }
}  // namespace tc
}  // namespace gnas
#endif  // GNAS_EMUL

// Functional model of the Blackwell pieces the tensor-core RHN kernel uses -- TEST INFRASTRUCTURE ONLY
// (included by cuda_emul.h): mbarrier (arrive / expect_tx / complete_tx / parity wait), 1-D bulk copies,
// named barriers, tensor memory (128 lanes x 512 columns per CTA), tcgen05.mma kind::tf32 with K-major
// no-swizzle shared-memory descriptors (8-row x 16-byte core matrices, LBO between core matrices along K,
// SBO between 8-row groups), tcgen05.ld 32x32b and tcgen05.commit.  MMAs execute synchronously when issued.
// The TMEM row -> lane map for M = 64 (rows 16q..16q+15 in lanes 32q..32q+15) was measured on a B200
// (tools/umma_probe.cu, profiles/r02_umma_probe.txt).  Checks kernel LOGIC (layouts, descriptors, pipeline
// hand-offs); rounding follows the hardware only loosely (operands truncated to tf32, fp32 accumulation).
#pragma once

namespace emul {

// ---- mbarrier: bits 0..19 pending arrivals, 20..39 arrival count of a phase, 40..62 transaction bytes
// (biased by 2^22), bit 63 phase parity
struct MbarView {
  uint64_t* w;
  static constexpr uint64_t kBias = 1ull << 22;
  uint32_t pending() const { return (uint32_t)(*w & 0xFFFFF); }
  uint32_t count() const { return (uint32_t)((*w >> 20) & 0xFFFFF); }
  int64_t tx() const { return (int64_t)((*w >> 40) & 0x7FFFFF) - (int64_t)kBias; }
  uint32_t phase() const { return (uint32_t)(*w >> 63); }
  void set(uint32_t pend, uint32_t cnt, int64_t tx_, uint32_t ph) {
    *w = (uint64_t)(pend & 0xFFFFF) | ((uint64_t)(cnt & 0xFFFFF) << 20) | ((uint64_t)((tx_ + (int64_t)kBias) & 0x7FFFFF) << 40) |
         ((uint64_t)ph << 63);
  }
  void settle() { if (pending() == 0 && tx() == 0) set(count(), count(), 0, phase() ^ 1u); }
};
inline void mbar_init(uint64_t* b, uint32_t count) { MbarView v{b}; v.set(count, count, 0, 0); }
inline void mbar_arrive(uint64_t* b, int64_t expect = 0) {
  MbarView v{b};
  if (v.pending() == 0) { fprintf(stderr, "emul: mbarrier over-arrival\n"); abort(); }
  v.set(v.pending() - 1, v.count(), v.tx() + expect, v.phase());
  v.settle();
}
inline void mbar_complete_tx(uint64_t* b, int64_t bytes) {
  MbarView v{b};
  v.set(v.pending(), v.count(), v.tx() - bytes, v.phase());
  v.settle();
}
inline bool mbar_test(uint64_t* b, uint32_t parity) { MbarView v{b}; return v.phase() != (parity & 1u); }

inline void named_bar_sync(int id, int nthreads) {
  NamedBarrier& nb = cur->block->named[id & 15];
  const unsigned g = nb.gen;
  if (++nb.arrived >= nthreads) { nb.arrived = 0; nb.gen++; }
  else while (nb.gen == g) yield();
}

// ---- shared-memory window: 32-bit addresses are byte offsets into the CTA's dynamic shared memory
inline uint32_t smem_offset(const void* p) {
  const ptrdiff_t o = reinterpret_cast<const unsigned char*>(p) - cur->block->dyn_smem;
  if (o < 0 || o >= (1 << 18)) { fprintf(stderr, "emul: pointer outside dynamic shared memory used as an operand\n"); abort(); }
  return (uint32_t)o;
}

// ---- tensor memory
inline float* tmem_base() {
  Block* b = cur->block;
  if (b->tmem.empty()) b->tmem.assign(128 * 512, 0.f);
  return b->tmem.data();
}
inline float& tmem_at(uint32_t lane, uint32_t col) {
  if (lane >= 128 || col >= 512) { fprintf(stderr, "emul: TMEM access out of range (lane %u col %u)\n", lane, col); abort(); }
  return tmem_base()[lane * 512 + col];
}

inline float tf32_trunc(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xFFFFE000u; memcpy(&x, &u, 4); return x; }
inline uint32_t tf32_rna(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xFFFFE000u; return u; }

// D[M x N] (+)= A[M x 8] * B[N x 8]^T, both operands K-major in the canonical no-swizzle layout
inline void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  const int N = (int)((idesc >> 17) & 0x3F) << 3, M = (int)((idesc >> 24) & 0x1F) << 4;
  if ((M != 64 && M != 128) || N < 8 || N > 256 || (N & 7)) { fprintf(stderr, "emul: unsupported MMA shape %d x %d\n", M, N); abort(); }
  auto field = [](uint64_t d, int lo) { return (uint32_t)((d >> lo) & 0x3FFF) << 4; };
  const unsigned char* sm = cur->block->dyn_smem;
  const uint32_t a0 = field(adesc, 0), a_lbo = field(adesc, 16), a_sbo = field(adesc, 32);
  const uint32_t b0 = field(bdesc, 0), b_lbo = field(bdesc, 16), b_sbo = field(bdesc, 32);
  auto at = [&](uint32_t base, uint32_t lbo, uint32_t sbo, int row, int k) {
    float v;
    memcpy(&v, sm + base + (uint32_t)(row >> 3) * sbo + (uint32_t)(k >> 2) * lbo + (uint32_t)(row & 7) * 16 + (uint32_t)(k & 3) * 4, 4);
    return tf32_trunc(v);
  };
  const uint32_t col0 = tmem_d & 0xFFFF, lane0 = tmem_d >> 16;
  for (int m = 0; m < M; ++m) {
    const uint32_t lane = lane0 + (M == 128 ? (uint32_t)m : (uint32_t)(32 * (m >> 4) + (m & 15)));
    for (int n = 0; n < N; ++n) {
      double s = 0.0;
      for (int k = 0; k < 8; ++k) s += (double)at(a0, a_lbo, a_sbo, m, k) * (double)at(b0, b_lbo, b_sbo, n, k);
      // the tensor core truncates (toward zero) when it adds into the fp32 accumulator: measured on B200 through the
      // RHN recurrence (2.1e-5 at T = 42 with one 64-step chain) and reproduced by exactly this model
      float& d = tmem_at(lane, col0 + (uint32_t)n);
      const double r = (accumulate ? (double)d : 0.0) + s;
      float t = (float)r;
      if (fabs((double)t) > fabs(r)) t = nextafterf(t, 0.f);
      d = t;
    }
  }
}

// tcgen05.ld.32x32b.x32: lane l of the warp reads TMEM lane (taddr.lane + l), 32 consecutive columns
inline void tmem_ld32(uint32_t taddr, float* v) {
  const uint32_t lane = (taddr >> 16) + (uint32_t)(cur->lin & 31), col = taddr & 0xFFFF;
  const uint32_t quarter = (uint32_t)((cur->lin >> 5) & 3);
  if ((taddr >> 16) != 32 * quarter) { fprintf(stderr, "emul: warp %d reads TMEM lanes of another quarter\n", cur->lin >> 5); abort(); }
  for (int i = 0; i < 32; ++i) v[i] = tmem_at(lane, col + (uint32_t)i);
}

}  // namespace emul

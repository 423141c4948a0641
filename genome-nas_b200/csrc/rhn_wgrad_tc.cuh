// RHN weight gradients on the 5th-generation tensor cores (3xTF32, fp32 accumulate in TMEM).
//   dWs[i][n][k'] = sum_{t, j<=i, m} (hm .* s_j(t))[n][m] * dCH_{i,j}(t)[k'][m]       (model_search.py:54-97, backward)
//   dW0[n][k']    = sum_{t, m}       [xm(t) ; hm .* h(t-1)][n][m] * dCH_0(t)[k'][m]   (model_discCNN.py:226-268, backward)
// The forward keeps every state feature-major ([H][BP], batch contiguous), so both operands of the GEMM are K-major
// as they lie in memory: A = 128 feature rows x 32 batch columns, B = 128 pre-activation rows x 32 batch columns per
// pipeline stage.  One CTA owns a 128 x 128 tile of one edge (i, j) (or of W0 for a range of t) and adds it to
// the gradient with red.global.add.v4.f32; split-K over edges keeps the per-accumulator chain short (112 MMAs).
#pragma once
#include "tc_common.cuh"

#ifndef GNAS_EMUL
namespace gnas {
namespace tc {

struct RhnWgTc {
  int T, B, BP, H, is_w0, tsplit;
  const float* ST; const float* CHT; const float* C0T; const float* HT; const float* h0T;
  const float* xmT; const float* hmT;
  float* dW0; float* dWs;
};

constexpr int kWgStages = 3;
constexpr int kWgOp = 8 * 128 * 4;                 // floats of one operand half: [8 k-chunks][128 rows][4]
constexpr int kWgStage = 4 * kWgOp;                // A_hi | A_lo | B_hi | B_lo
constexpr int kWgLoaders = kTcThreads - 32;        // warps 1..7
constexpr size_t kWgSmem = sizeof(float) * (size_t)kWgStage * kWgStages + 128;

static inline bool rhn_wgrad_tc_supported(int B, int BP, int H) { return H % 128 == 0 && BP % 32 == 0 && B % 4 == 0; }

__global__ void __launch_bounds__(kTcThreads, 1) k_rhn_wgrad_tc(const __grid_constant__ RhnWgTc P) {
  const int H = P.H, BP = P.BP, B = P.B;
  const long long HB = (long long)H * BP;
  const int n0 = blockIdx.x * 128, k0 = blockIdx.y * 128;
  const int nh = BP / 32;                                   // stages per time step
  int iw = 0, jj = 0, tb = 0, te = P.T;
  if (P.is_w0) {
    const int tper = (P.T + P.tsplit - 1) / P.tsplit;
    tb = blockIdx.z * tper; te = min(P.T, tb + tper);
    if (tb >= te) return;
  } else {
    int e = blockIdx.z;
    while (e > iw) { e -= iw + 1; ++iw; }
    jj = e;
  }
  const int nstage = (te - tb) * nh;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  extern __shared__ __align__(128) unsigned char wg_smem_raw[];
  float* sm = reinterpret_cast<float*>(wg_smem_raw);
  __shared__ __align__(8) uint64_t bar_full[kWgStages], bar_empty[kWgStages], bar_done;
  __shared__ uint32_t tmem_base_s;
  constexpr uint32_t LBO = 128 * 16, SBO = 128;
  constexpr int TM_COLS = 512;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
#pragma unroll
    for (int i = 0; i < kWgStages; ++i) { mbar_init(&bar_full[i], kWgLoaders); mbar_init(&bar_empty[i], 1); }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0) {
    // ---- MMA issuer (whole warp loops, one elected lane issues)
    const uint32_t idesc = make_idesc(128);
    const uint64_t d0 = make_desc(smem_u32(sm), LBO, SBO);
    int cnt = 0;
    for (int st = 0; st < nstage; ++st) {
      const int s = st % kWgStages;
      mbar_wait(&bar_full[s], (st / kWgStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t base = d0 + (uint64_t)(((uint32_t)s * kWgStage * 4) >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t koff = (uint64_t)((ks * 2 * LBO) >> 4);
          const uint64_t ah = base + koff, al = ah + (uint64_t)((kWgOp * 4) >> 4);
          const uint64_t bh = base + koff + (uint64_t)((2 * kWgOp * 4) >> 4), bl = bh + (uint64_t)((kWgOp * 4) >> 4);
          const int unit = cnt + ks;
          mma_tf32(tmem_d + (unit % 3) * 128, ah, bh, idesc, (uint32_t)(unit >= 3));   // hi*hi: accumulators 0..2 in turn
          mma_tf32(tmem_d + 3 * 128, ah, bl, idesc, (uint32_t)(unit > 0));              // cross terms: accumulator 3
          mma_tf32(tmem_d + 3 * 128, al, bh, idesc, 1);
        }
        mma_commit(&bar_empty[s]);
      }
      cnt += 4;
      __syncwarp();
    }
    if (elect_one()) mma_commit(&bar_done);
    __syncwarp();
  } else {
    // ---- loaders: global float4 (batch-contiguous) -> mask -> tf32 hi / lo -> canonical K-major core matrices
    const int lt = tid - 32;
    constexpr int NIT = 2048, PER = (NIT + kWgLoaders - 1) / kWgLoaders;   // items per stage, per thread
    for (int st = 0; st < nstage; ++st) {
      const int s = st % kWgStages;
      if (st >= kWgStages) mbar_wait(&bar_empty[s], ((st / kWgStages) - 1) & 1);
      const int t = tb + st / nh, m0 = (st % nh) * 32;
      float* stage = sm + (size_t)s * kWgStage;
      float4 v[PER], mk[PER];
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int idx = lt + i * kWgLoaders;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f); mk[i] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (idx >= NIT) continue;
        const int op = idx >> 10, q = idx & 1023, rb = q >> 5, w = q & 31;
        const int r = (rb & 7) * 16 + (w & 15), j = 2 * (rb >> 3) + (w >> 4);
        const int m = m0 + 4 * j;
        if (m >= B) continue;
        if (op == 0) {
          const int n = n0 + r;
          if (P.is_w0) {
            if (n < H) v[i] = *reinterpret_cast<const float4*>(P.xmT + (size_t)t * HB + (size_t)n * BP + m);
            else {
              const size_t o = (size_t)(n - H) * BP + m;
              v[i] = *reinterpret_cast<const float4*>((t > 0 ? P.HT + (size_t)(t - 1) * HB : P.h0T) + o);
              mk[i] = *reinterpret_cast<const float4*>(P.hmT + o);
            }
          } else {
            const size_t o = (size_t)n * BP + m;
            v[i] = *reinterpret_cast<const float4*>(P.ST + ((size_t)t * 9 + jj) * HB + o);
            mk[i] = *reinterpret_cast<const float4*>(P.hmT + o);
          }
        } else {
          const float* D = P.is_w0 ? P.C0T + ((size_t)t * 2) * HB : P.CHT + (((size_t)t * 36 + iw * (iw + 1) / 2 + jj) * 2) * HB;
          v[i] = *reinterpret_cast<const float4*>(D + (size_t)(k0 + r) * BP + m);
        }
      }
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int idx = lt + i * kWgLoaders;
        if (idx >= NIT) continue;
        const int op = idx >> 10, q = idx & 1023, rb = q >> 5, w = q & 31;
        const int r = (rb & 7) * 16 + (w & 15), j = 2 * (rb >> 3) + (w >> 4);
        const float a0 = v[i].x * mk[i].x, a1 = v[i].y * mk[i].y, a2 = v[i].z * mk[i].z, a3 = v[i].w * mk[i].w;
        const float h0 = __uint_as_float(to_tf32(a0)), h1 = __uint_as_float(to_tf32(a1));
        const float h2 = __uint_as_float(to_tf32(a2)), h3 = __uint_as_float(to_tf32(a3));
        float* dst = stage + (size_t)op * 2 * kWgOp + ((size_t)j * 128 + r) * 4;
        *reinterpret_cast<float4*>(dst) = make_float4(h0, h1, h2, h3);
        *reinterpret_cast<float4*>(dst + kWgOp) = make_float4(__uint_as_float(to_tf32(a0 - h0)), __uint_as_float(to_tf32(a1 - h1)),
                                                              __uint_as_float(to_tf32(a2 - h2)), __uint_as_float(to_tf32(a3 - h3)));
      }
      fence_async_smem();
      mbar_arrive(&bar_full[s]);
    }
  }
  __syncwarp();

  // ---- epilogue: TMEM -> registers -> red.add into the gradient
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  {
    const int row = 32 * (warp & 3) + lane;
    const int n = n0 + row;
    const int ldw = 2 * H;
    float* dW = P.is_w0 ? P.dW0 : P.dWs + (size_t)iw * H * ldw;
    float* orow = dW + (size_t)n * ldw + k0;
#pragma unroll 1
    for (int c0 = (warp >> 2) * 64; c0 < (warp >> 2) * 64 + 64; c0 += 32) {
      float v[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0;
      tmem_ld32(taddr, v);
#pragma unroll
      for (int a = 1; a < 4; ++a) {
        float t[32];
        tmem_ld32(taddr + a * 128, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += t[i];
      }
#pragma unroll
      for (int i = 0; i < 32; i += 4) red_add_v4(orow + c0 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TM_COLS) : "memory");
  }
}

static void launch_rhn_wgrad_tc(RhnWgTc& w, cudaStream_t st) {
  static int prepared[kMaxDevices] = {0};
  if (need_prepare(prepared, (int)kWgSmem)) cudaFuncSetAttribute(k_rhn_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmem);
  const int H = w.H;
  w.is_w0 = 0; w.tsplit = 1;
  prof_before("k_rhn_wgrad_tc", "", st);
  k_rhn_wgrad_tc<<<dim3(H / 128, 2 * H / 128, 36), dim3(kTcThreads), kWgSmem, st>>>(w);
  prof_after(st);
  w.is_w0 = 1; w.tsplit = w.T >= 3 ? 3 : 1;
  prof_before("k_rhn_wgrad_tc", "", st);
  k_rhn_wgrad_tc<<<dim3(2 * H / 128, 2 * H / 128, w.tsplit), dim3(kTcThreads), kWgSmem, st>>>(w);
  prof_after(st);
}

}  // namespace tc
}  // namespace gnas
#endif  // !GNAS_EMUL

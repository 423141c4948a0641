// Recurrent highway layer of the search cell (DARTSCell.forward + DARTSCellSearch.cell).
// Replaces darts_tools/model_discCNN.py:226-268 and model_search.py:54-97 behind include/gnas.h.
//
// Forward: ONE persistent cooperative kernel.  CTA c owns hidden columns [4c, 4c+4): its slices of
// W0 (2H x 8) and of all eight Ws[i] (H x 8 each) stay in shared memory for the whole sequence
// (H = 512: 160 KB).  BatchNorm over the batch is per column, hence CTA-local; each new state is
// published (masked) to a ring in global memory and a grid barrier separates the 9 stages of a time
// step.  When s_j appears it is multiplied at once with the slices of Ws[j..7] ("multiply on
// arrival"): same FLOPs as the reference's per-step re-multiplication of all states, no concat.
// All internal state is feature-major ([H][B]) so that a CTA's columns are contiguous rows.
//
// Backward: staged launches per (time step, node) -- elementwise/BN kernel + split-K GEMM against
// Ws[i]^T -- and, off the critical path, one batched GEMM pass for dW0 / dWs over all time steps.
#include "common.cuh"
#include <cstdlib>
#include "tc_common.cuh"
#include "rhn_wgrad_tc.cuh"
#include "../../include/gnas.h"

namespace gnas {

constexpr int kNC = 4;        // hidden columns per CTA in the persistent forward kernel
constexpr int kKCH = 32;      // K rows per shared-memory chunk
constexpr int kMaxBP = 64;    // padded batch rows supported by the persistent kernel
constexpr int kRhnThreads = 512;   // persistent forward CTA: 16 warps hide the shared-memory latency of the register-tiled product

#ifdef GNAS_EMUL
__device__ __forceinline__ void cp_async16(void* dst, const void* src) { memcpy(dst, src, 16); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N> __device__ __forceinline__ void cp_async_wait() {}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) { return emul_ld_acquire(p); }
__device__ __forceinline__ int ld_volatile_i32(const int* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ float ld_cg(const float* p) { return *p; }
__device__ __forceinline__ float4 ld_cg4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ long long gnas_clock() { return 0; }
#else
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_volatile_i32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ float ld_cg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ long long gnas_clock() { return clock64(); }
#endif

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// activation k in {0: tanh, 1: identity, 2: sigmoid, 3: relu}  (PRIMITIVES_rnn[1:], genotypes.py:42-48)
__device__ __forceinline__ float act_f(int k, float h) {
  switch (k) { case 0: return tanhf(h); case 1: return h; case 2: return sigmoidf_(h); default: return h > 0.f ? h : 0.f; }
}
__device__ __forceinline__ float act_df(int k, float h) {
  switch (k) {
    case 0: { const float t = tanhf(h); return 1.f - t * t; }
    case 1: return 1.f;
    case 2: { const float s = sigmoidf_(h); return s * (1.f - s); }
    default: return h > 0.f ? 1.f : 0.f;
  }
}

struct RhnTable {             // per edge p = i(i+1)/2 + j: index into probs for each activation, -1 if disabled
  short pidx[GNAS_RNN_EDGES][4];
};

}  // namespace gnas
#include "rhn_tc.cuh"
namespace gnas {

// The tensor-core forward (rhn_tc.cuh) is the default wherever its shape constraints hold; GNAS_RHN_TC=0 (or
// GNAS_TC=0 on the GPU) selects the fp32 CUDA-core kernel below, which also serves the other shapes.
static inline bool rhn_tc_wanted() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GNAS_RHN_TC"); v = e ? atoi(e) : 1; }
#ifndef GNAS_EMUL
  return v != 0 && tc_enabled();
#else
  return v != 0;
#endif
}

struct RhnPlan {
  int BP;
  long long xmT, xkT, hmT, h0T, ring, ST, CHT, C0T, HT, stat, bar, tc, fwd_total;   // forward workspace (floats)
  long long dsT, dhT, dxT, doT, dpacc, WsT, W0T, bwd_zero, bwd_total;           // backward scratch
};

static RhnPlan rhn_plan(const GnasRhnDesc* d) {
  RhnPlan p;
  const long long T = d->T, H = d->H;
  p.BP = (d->B + 3) / 4 * 4;
  const long long HB = H * p.BP;
  long long o = 0;
  p.xmT = o; o += T * HB;          // (x * x_mask)^T
  p.xkT = o; o += HB;              // x_mask^T
  p.hmT = o; o += HB;              // h_mask^T
  p.h0T = o; o += HB;
  p.ring = o; o += 10 * HB;        // masked s_0..s_8, masked h
  p.ST = o; o += T * 9 * HB;       // unmasked states
  p.CHT = o; o += T * 36 * 2 * HB; // pre-activations c|h per edge (overwritten by dc|dh in backward)
  p.C0T = o; o += T * 2 * HB;
  p.HT = o; o += T * HB;           // h_t^T
  p.stat = o; o += T * 9 * 2 * H;
  p.bar = o; o += 256;             // barrier counter + abort flag (+ debug timing marks)
  p.tc = o;                        // operands of the tensor-core forward: x, state ring, packed weights (hi | lo)
  if (H % rtc::KCH == 0 && d->B <= rtc::BPT) o += rtc::plan(d->T, d->H).total;
  p.fwd_total = o;
  o = 0;
  p.dpacc = o; o += 2 * 36 * 4;    // doubles
  p.dsT = o; o += 9 * HB;
  p.dhT = o; o += 2 * HB;          // ping-pong
  p.dxT = o; o += T * HB;
  p.bwd_zero = o;                  // everything above is zero-initialised per call
  p.WsT = o; o += 8 * H * 2 * H;   // Ws[i]^T  [2H][H]  (so that GEMM tiles are straight 16-byte copies)
  p.W0T = o; o += 2 * H * 2 * H;   // W0^T     [2H][2H]
  p.bwd_total = o;
  return p;
}

static void rhn_table(const GnasRhnDesc* d, RhnTable& tb) {
  // comp_aux.py:127-233: probability row = edge index - #fully disabled edges so far; column = #enabled before k
  int disc = 0;
  for (int r = 0; r < GNAS_RNN_EDGES; ++r) {
    bool any = false;
    for (int k = 0; k < GNAS_RNN_PRIMS; ++k) any = any || d->sw[r][k];
    if (!any) ++disc;
    for (int k = 1; k < GNAS_RNN_PRIMS; ++k) {
      int col = 0;
      for (int q = 0; q < k; ++q) col += d->sw[r][q] ? 1 : 0;
      tb.pidx[r][k - 1] = d->sw[r][k] ? (short)((r - disc) * d->n_prims + col) : (short)-1;
    }
  }
}

// ------------------------------------------------------------------ prologue / epilogue transposes
struct RhnPro {
  int T, B, BP, H;
  const float* x; const float* h0; const float* xm; const float* hm;
  float* xmT; float* xkT; float* hmT; float* h0T; float* ring9;
};
static __global__ void __launch_bounds__(256) k_rhn_prologue(const __grid_constant__ RhnPro P) {
  const long long HB = (long long)P.H * P.BP;
  const long long total = (long long)(P.T + 3) * HB;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int slab = (int)(i / HB);
    const int r = (int)(i - (long long)slab * HB);
    const int n = r / P.BP, m = r - n * P.BP;
    const bool real = m < P.B;
    const float xk = real ? (P.xm ? P.xm[m * P.H + n] : 1.f) : 0.f;
    const float hk = real ? (P.hm ? P.hm[m * P.H + n] : 1.f) : 0.f;
    if (slab < P.T) P.xmT[i] = real ? P.x[((long long)slab * P.B + m) * P.H + n] * xk : 0.f;
    else if (slab == P.T) P.xkT[r] = xk;
    else if (slab == P.T + 1) P.hmT[r] = hk;
    else {
      const float h = real ? P.h0[m * P.H + n] : 0.f;
      P.h0T[r] = h;
      P.ring9[r] = h * hk;
    }
  }
}

// dst[t][m][n] = src[t][n][m]   ([H][BP] -> [B][H])
static __global__ void __launch_bounds__(256) k_rhn_untranspose(const float* src, float* dst, int T, int B, int BP, int H) {
  const long long total = (long long)T * B * H;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % H);
    const int m = (int)((i / H) % B);
    const long long t = i / ((long long)H * B);
    dst[i] = src[(t * H + n) * BP + m];
  }
}

// ------------------------------------------------------------------ persistent forward kernel
struct RhnFwd {
  int T, B, BP, H, training;
  float momentum;
  const float* xmT; const float* hmT; const float* h0T;
  const float* W0; const float* Ws; const float* probs;
  float* ring; float* ST; float* CHT; float* C0T; float* HT; float* stat;
  float* out; float* running;
  unsigned* bar; int* abort_flag;
  RhnTable tb;
};

// Grid barrier split into arrive / wait so that work that does not gate the next stage (stashing the
// pre-activations for backward, the contributions to later nodes) overlaps the wait.
__device__ __forceinline__ void grid_arrive(unsigned* bar) {
  __syncthreads();                       // every thread's ring writes are issued
  if (threadIdx.x == 0) { __threadfence(); atomicAdd(bar, 1u); }
}
__device__ __forceinline__ bool grid_wait(unsigned* bar, int* abort_flag, unsigned target, int* s_abort) {
  if (threadIdx.x == 0) {
    const long long t0 = gnas_clock();
    int ab = 0;
    while (ld_acquire_u32(bar) < target) {
      GNAS_SPIN_PAUSE();
      if (ld_volatile_i32(abort_flag)) { ab = 1; break; }
      if (gnas_clock() - t0 > 4000000000LL) { atomicAdd(abort_flag, 1); ab = 1; break; }   // ~2 s: never hang the GPU
    }
    __threadfence();
    *s_abort = ab;
  }
  __syncthreads();
  return *s_abort != 0;
}

// K-split factor: power of two, fills the CTA, and keeps KS * N <= 64 columns of the partial buffer
__device__ __forceinline__ int rhn_ks(int BP, int N) {
  int a = kRhnThreads / ((BP / 4) * (N / 4));
  const int b = 128 / N;                 // partial buffer = the ring: kStages*kKCH*BP = 128*BP floats
  a = a < b ? a : b;
  return a >= 16 ? 16 : (a >= 8 ? 8 : (a >= 4 ? 4 : (a >= 2 ? 2 : 1)));
}

constexpr int kStages = 4;    // cp.async ring depth of the streamed operand

// one kKCH-row chunk of the register-tiled product: 4 batch rows x 4 columns per thread, every KS-th k.
// Compile-time trip count + explicit one-deep software pipeline so the shared loads of step i+1 are in
// flight while the 16 FMAs of step i issue (only 2 warps per scheduler are resident).
template <int KS>
__device__ __forceinline__ void rhn_chunk(const float* __restrict__ as, const float* __restrict__ wk, int BP, float (&acc)[4][4]) {
  constexpr int NK = kKCH / KS;
  float4 a[2], w[2];
  a[0] = *reinterpret_cast<const float4*>(as);
  w[0] = *reinterpret_cast<const float4*>(wk);
#pragma unroll
  for (int i = 0; i < NK; ++i) {
    if (i + 1 < NK) {
      a[(i + 1) & 1] = *reinterpret_cast<const float4*>(as + (size_t)(i + 1) * KS * BP);
      w[(i + 1) & 1] = *reinterpret_cast<const float4*>(wk + (i + 1) * KS * 8);
    }
    const float4 x = a[i & 1], y = w[i & 1];
    acc[0][0] = fmaf(x.x, y.x, acc[0][0]); acc[0][1] = fmaf(x.x, y.y, acc[0][1]); acc[0][2] = fmaf(x.x, y.z, acc[0][2]); acc[0][3] = fmaf(x.x, y.w, acc[0][3]);
    acc[1][0] = fmaf(x.y, y.x, acc[1][0]); acc[1][1] = fmaf(x.y, y.y, acc[1][1]); acc[1][2] = fmaf(x.y, y.z, acc[1][2]); acc[1][3] = fmaf(x.y, y.w, acc[1][3]);
    acc[2][0] = fmaf(x.z, y.x, acc[2][0]); acc[2][1] = fmaf(x.z, y.y, acc[2][1]); acc[2][2] = fmaf(x.z, y.z, acc[2][2]); acc[2][3] = fmaf(x.z, y.w, acc[2][3]);
    acc[3][0] = fmaf(x.w, y.x, acc[3][0]); acc[3][1] = fmaf(x.w, y.y, acc[3][1]); acc[3][2] = fmaf(x.w, y.z, acc[3][2]); acc[3][3] = fmaf(x.w, y.w, acc[3][3]);
  }
}

// C[m][n] partial = sum_k A[k][m] * W[k][n]; the W slices are resident in shared memory, A ([K][BP], already
// masked) is streamed from L2 in kKCH-row chunks.  Measured on B200 (clock64 marks in the kernel): with a
// 4-deep cp.async ring the loop ran at ~850 cycles per 8 KB chunk whatever the math -- an L2 round trip of
// ~2500 cycles divided by 3 chunks in flight.  Shared memory is full (weights), so the in-flight window lives
// in REGISTERS: every thread keeps its 16-byte share of the next kPre chunks (ld.global.cg, 8 x 8 KB in
// flight per SM), and copies the oldest into a double-buffered shared tile just before it is needed.
// One __syncthreads per chunk.  The K-split partials are left in `part`, which aliases the tiles.
constexpr int kPre = 8;
template <int KS>
__device__ __forceinline__ void rhn_gemm_t(const float* __restrict__ A0, const float* __restrict__ A1, int K0, int K,
                                           int BP, int N, const float* wbase, int wstride_i, bool stage0,
                                           float* abuf, float* part) {
  const int tid = threadIdx.x;
  const int nrg = BP / 4, ncg = N / 4;
  const int rg = tid % nrg, cg = (tid / nrg) % ncg, ks = tid / (nrg * ncg);
  const bool active = ks < KS;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  // W column group: stage 0 -> W0s[k][8]; stage j -> Wss[i][k][8] with i = cg / 2
  const float* wk = (stage0 ? wbase + cg * 4 : wbase + (size_t)(cg >> 1) * wstride_i + (cg & 1) * 4) + ks * 8;
  const int nch = K / kKCH;
  const int chunk_floats = kKCH * BP;
  const bool mine = 4 * tid < chunk_floats;          // this thread's 16-byte share of a chunk (BP = 64: everyone)
  auto fetch = [&](int c) -> float4 {
    if (c < nch && mine) {
      const int k0 = c * kKCH;
      const float* src = (k0 < K0) ? A0 + (size_t)k0 * BP : A1 + (size_t)(k0 - K0) * BP;
      return ld_cg4(src + 4 * tid);
    }
    return make_float4(0.f, 0.f, 0.f, 0.f);
  };
  float4 r[kPre];
#pragma unroll
  for (int u = 0; u < kPre; ++u) r[u] = fetch(u);
  if (mine) *reinterpret_cast<float4*>(abuf + 4 * tid) = r[0];
  __syncthreads();
  const float* as0 = abuf + rg * 4 + ks * BP;
  for (int c0 = 0; c0 < nch; c0 += kPre) {
#pragma unroll
    for (int u = 0; u < kPre; ++u) {
      const int c = c0 + u;
      if (c < nch) {
        // chunk c+1 (in registers since kPre-1 iterations) -> the tile chunk c-1 was read from
        if (c + 1 < nch && mine) *reinterpret_cast<float4*>(abuf + ((c + 1) & 1) * chunk_floats + 4 * tid) = r[(u + 1) % kPre];
        r[u] = fetch(c + kPre);
        if (active) rhn_chunk<KS>(as0 + (c & 1) * chunk_floats, wk, BP, acc);
        wk += kKCH * 8;
        __syncthreads();
      }
    }
  }
  // partials are stored column-major ([ks][n][BP], 4 consecutive batch rows per 128-bit store): conflict-free
  // for these stores and for the batch-row-fastest reads of the elementwise phase
  if (active) {
#pragma unroll
    for (int b = 0; b < 4; ++b)
      *reinterpret_cast<float4*>(part + ((size_t)(ks * N + cg * 4 + b)) * BP + rg * 4) = make_float4(acc[0][b], acc[1][b], acc[2][b], acc[3][b]);
  }
  __syncthreads();
}

__device__ __forceinline__ void rhn_gemm(const float* __restrict__ A0, const float* __restrict__ A1, int K0, int K,
                                         int BP, int N, const float* wbase, int wstride_i, bool stage0,
                                         float* abuf, float* part) {
  switch (rhn_ks(BP, N)) {
    case 1: rhn_gemm_t<1>(A0, A1, K0, K, BP, N, wbase, wstride_i, stage0, abuf, part); break;
    case 2: rhn_gemm_t<2>(A0, A1, K0, K, BP, N, wbase, wstride_i, stage0, abuf, part); break;
    case 4: rhn_gemm_t<4>(A0, A1, K0, K, BP, N, wbase, wstride_i, stage0, abuf, part); break;
    case 8: rhn_gemm_t<8>(A0, A1, K0, K, BP, N, wbase, wstride_i, stage0, abuf, part); break;
    default: rhn_gemm_t<16>(A0, A1, K0, K, BP, N, wbase, wstride_i, stage0, abuf, part); break;
  }
}

__global__ void __launch_bounds__(kRhnThreads, 1) k_rhn_forward(const __grid_constant__ RhnFwd P) {
  const int H = P.H, B = P.B, BP = P.BP, T = P.T;
  const int n0 = blockIdx.x * kNC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long HB = (long long)H * BP;
  GNAS_DYN_SMEM(float, smem);
  float* W0s = smem;                          // [2H][8]
  float* Wss = W0s + (size_t)2 * H * 8;       // [8][H][8]
  float* abuf = Wss + (size_t)8 * H * 8;      // [2][kKCH][BP] operand tiles; the region is 128*BP floats because
  float* part = abuf;                         // ... the K-split partials [KS][N][BP] (<= 128*BP floats) alias it
  float* accn = abuf + kStages * kKCH * BP;   // [8][BP][4]  pending node accumulators
  float* sloc = accn + 8 * BP * 4;            // [9][BP][4]  own columns of s_0..s_8
  float* hloc = sloc + 9 * BP * 4;            // [BP][4]     own columns of h_{t-1}
  float* hml = hloc + BP * 4;                 // [BP][4]     own columns of the hidden mask
  float* tmp0 = hml + BP * 4;                 // [BP][4]     raw s_0 before BN
  float* pp = tmp0 + BP * 4;                  // [36][4] probabilities, [36] q
  float* pq = pp + 36 * 4;
  float* run = pq + 36;                       // [2][4] running mean | var of own columns
  __shared__ int s_abort;
  // ---- one-time setup: weight slices, masks, probability table
  for (int i = tid; i < 2 * H * 8; i += kRhnThreads) {
    const int k = i >> 3, c = i & 7;
    W0s[i] = P.W0[(size_t)k * 2 * H + (c < 4 ? n0 + c : H + n0 + c - 4)];
  }
  for (int i = tid; i < 8 * H * 8; i += kRhnThreads) {
    const int c = i & 7, k = (i >> 3) % H, w = i / (8 * H);
    Wss[i] = P.Ws[((size_t)w * H + k) * 2 * H + (c < 4 ? n0 + c : H + n0 + c - 4)];
  }
  for (int i = tid; i < BP * 4; i += kRhnThreads) {
    const int m = i >> 2, c = i & 3;
    hml[i] = P.hmT[(size_t)(n0 + c) * BP + m];
    hloc[i] = P.h0T[(size_t)(n0 + c) * BP + m];
  }
  for (int i = tid; i < 8 * BP * 4; i += kRhnThreads) accn[i] = 0.f;
  for (int i = tid; i < 9 * BP * 4; i += kRhnThreads) sloc[i] = 0.f;
  if (tid < 36) {
    float q = 0.f;
    for (int k = 0; k < 4; ++k) {
      const int ix = P.tb.pidx[tid][k];
      const float v = ix >= 0 ? P.probs[ix] : 0.f;
      pp[tid * 4 + k] = v; q += v;
    }
    pq[tid] = q;
  }
  if (tid < 8) run[tid] = P.running ? P.running[(tid >> 2) * H + n0 + (tid & 3)] : (tid < 4 ? 0.f : 1.f);
  __syncthreads();
  unsigned bar_target = 0;
  const float invB = 1.f / (float)B;
#ifdef GNAS_RHN_TIMING
  long long* dbg = reinterpret_cast<long long*>(P.bar + 16);   // [stage 0..8][4 marks] of time step 1, CTA 0
#define GNAS_MARK(stage, k) do { if (blockIdx.x == 0 && t == 1 && tid == 0) dbg[(stage) * 4 + (k)] = clock64(); } while (0)
#else
#define GNAS_MARK(stage, k) do { } while (0)
#endif

  // BatchNorm over the batch for the 4 own columns of `src` ([BP][4]); warp w < 4 handles column w and
  // publishes the masked state to the ring (what the other CTAs wait for).
  auto bn_publish = [&](const float* src, int t, int stage) {
    if (warp < kNC) {
      const int col = warp;
      float v0 = lane < B ? src[lane * 4 + col] : 0.f;
      float v1 = lane + 32 < B ? src[(lane + 32) * 4 + col] : 0.f;
      float mean, rstd;
      if (P.training) {
        mean = warp_sum(v0 + v1) * invB;
        const float d0 = lane < B ? v0 - mean : 0.f, d1 = lane + 32 < B ? v1 - mean : 0.f;
        const float var = warp_sum(d0 * d0 + d1 * d1) * invB;
        rstd = 1.0f / sqrtf(var + kBnEps);
        if (lane == 0) {
          run[col] = (1.f - P.momentum) * run[col] + P.momentum * mean;
          run[4 + col] = (1.f - P.momentum) * run[4 + col] + P.momentum * (B > 1 ? var * (float)B / (float)(B - 1) : var);
        }
      } else {
        mean = run[col];
        rstd = 1.0f / sqrtf(run[4 + col] + kBnEps);
      }
      const int n = n0 + col;
      if (lane == 0) {
        float* st = P.stat + ((size_t)t * 9 + stage) * 2 * H;
        st[n] = mean; st[H + n] = rstd;
      }
      float* rgr = P.ring + (size_t)stage * HB + (size_t)n * BP;
      for (int m = lane; m < BP; m += 32) {
        const float v = m < B ? ((m < 32 ? v0 : v1) - mean) * rstd : 0.f;
        sloc[(stage * BP + m) * 4 + col] = v;
        rgr[m] = v * hml[m * 4 + col];
      }
    }
  };
  // unmasked state for backward (off the critical path)
  auto stash_state = [&](int t, int stage) {
    for (int i = tid; i < BP * 4; i += kRhnThreads) {
      const int col = i / BP, m = i - col * BP;
      P.ST[((size_t)t * 9 + stage) * HB + (size_t)(n0 + col) * BP + m] = sloc[(stage * BP + m) * 4 + col];
    }
  };
  // contribution of edge (i, j) to node i from c|h in `part`
  auto edge_item = [&](int t, int j, int irel, int m, int col, int N, int KS, bool store) {
    const int i = j + irel;
    const int p = i * (i + 1) / 2 + j;
    float c = 0.f, h = 0.f;
    for (int s = 0; s < KS; ++s) {
      c += part[(size_t)(s * N + irel * 8 + col) * BP + m];
      h += part[(size_t)(s * N + irel * 8 + 4 + col) * BP + m];
    }
    if (store) {
      float* chp = P.CHT + (((size_t)t * 36 + p) * 2) * HB + (size_t)(n0 + col) * BP + m;
      chp[0] = c; chp[HB] = h;
    }
    const float q = pq[p];
    const float p0 = pp[p * 4], p1 = pp[p * 4 + 1], p2 = pp[p * 4 + 2], p3 = pp[p * 4 + 3];
    if (p0 != 0.f || p1 != 0.f || p2 != 0.f || p3 != 0.f) {
      const float sj = sloc[(j * BP + m) * 4 + col];
      const float g = sigmoidf_(c);
      float F = 0.f;
      if (p0 != 0.f) F = fmaf(p0, tanhf(h), F);
      if (p1 != 0.f) F = fmaf(p1, h, F);
      if (p2 != 0.f) F = fmaf(p2, sigmoidf_(h), F);
      if (p3 != 0.f) F = fmaf(p3, h > 0.f ? h : 0.f, F);
      accn[(i * BP + m) * 4 + col] += q * sj + g * (F - q * sj);
    }
  };

  for (int t = 0; t < T; ++t) {
    // ---------------- stage 0: [c0|h0] = [x_t*xm | h*hm] W0 ; s_0 = BN(h + sig(c0)(tanh(h0) - h))
    GNAS_MARK(0, 0);
    rhn_gemm(P.xmT + (size_t)t * HB, P.ring + 9 * HB, H, 2 * H, BP, 8, W0s, 0, true, abuf, part);
    GNAS_MARK(0, 1);
    {
      const int KS = rhn_ks(BP, 8);
      for (int i = tid; i < B * 4; i += kRhnThreads) {
        const int col = i / B, m = i - col * B;
        float c = 0.f, h = 0.f;
        for (int s = 0; s < KS; ++s) { c += part[(size_t)(s * 8 + col) * BP + m]; h += part[(size_t)(s * 8 + 4 + col) * BP + m]; }
        P.C0T[((size_t)t * 2 + 0) * HB + (size_t)(n0 + col) * BP + m] = c;
        P.C0T[((size_t)t * 2 + 1) * HB + (size_t)(n0 + col) * BP + m] = h;
        const float hp = hloc[m * 4 + col];
        tmp0[m * 4 + col] = hp + sigmoidf_(c) * (tanhf(h) - hp);
      }
      __syncthreads();
      bn_publish(tmp0, t, 0);
    }
    GNAS_MARK(0, 2);
    grid_arrive(P.bar);
    stash_state(t, 0);
    bar_target += gridDim.x;
    if (grid_wait(P.bar, P.abort_flag, bar_target, &s_abort)) return;
    GNAS_MARK(0, 3);
    // ---------------- stages j = 0..7: multiply s_j with the slices of Ws[j..7]; node j -> s_{j+1} first
    for (int j = 0; j < 8; ++j) {
      const int N = 8 * (8 - j);
      GNAS_MARK(j + 1, 0);
      rhn_gemm(P.ring + (size_t)j * HB, nullptr, H, H, BP, N, Wss + (size_t)j * H * 8, H * 8, false, abuf, part);
      GNAS_MARK(j + 1, 1);
      const int KS = rhn_ks(BP, N);
      // critical: edge (j, j) completes node j
      for (int it = tid; it < 4 * B; it += kRhnThreads) edge_item(t, j, 0, it % B, it / B, N, KS, false);
      __syncthreads();
      bn_publish(accn + j * BP * 4, t, j + 1);
      const bool last = j == 7;
      if (last) {
        __syncthreads();
        // h_t = mean(s_1..s_8)  (model_search.py:96); its masked copy feeds stage 0 of the next step
        for (int i = tid; i < BP * 4; i += kRhnThreads) {
          const int m = i >> 2, col = i & 3;
          float a = 0.f;
#pragma unroll
          for (int s = 1; s <= 8; ++s) a += sloc[(s * BP + m) * 4 + col];
          a *= 0.125f;
          if (m >= B) a = 0.f;
          hloc[i] = a;
          P.ring[9 * HB + (size_t)(n0 + col) * BP + m] = a * hml[i];
        }
      }
      const bool need_barrier = !last || t + 1 < T;
      GNAS_MARK(j + 1, 2);
      if (need_barrier) grid_arrive(P.bar); else __syncthreads();
      // off the critical path: stash for backward, contributions of s_j to the later nodes, bookkeeping
      for (int it = tid; it < 4 * B; it += kRhnThreads) {        // pre-activations of edge (j, j)
        const int m = it % B, col = it / B;
        float c = 0.f, h = 0.f;
        for (int s = 0; s < KS; ++s) { c += part[(size_t)(s * N + col) * BP + m]; h += part[(size_t)(s * N + 4 + col) * BP + m]; }
        float* chp = P.CHT + (((size_t)t * 36 + j * (j + 1) / 2 + j) * 2) * HB + (size_t)(n0 + col) * BP + m;
        chp[0] = c; chp[HB] = h;
      }
      for (int it = tid; it < (7 - j) * 4 * B; it += kRhnThreads) {
        const int m = it % B, col = (it / B) & 3, irel = 1 + it / (4 * B);
        edge_item(t, j, irel, m, col, N, KS, true);
      }
      stash_state(t, j + 1);
      for (int i = tid; i < BP * 4; i += kRhnThreads) accn[j * BP * 4 + i] = 0.f;
      if (last) {
        for (int i = tid; i < BP * 4; i += kRhnThreads) {
          const int col = i / BP, m = i - col * BP;
          P.HT[(size_t)t * HB + (size_t)(n0 + col) * BP + m] = hloc[m * 4 + col];
        }
        for (int m = tid; m < B; m += kRhnThreads)
          *reinterpret_cast<float4*>(P.out + ((size_t)t * B + m) * H + n0) =
              make_float4(hloc[m * 4], hloc[m * 4 + 1], hloc[m * 4 + 2], hloc[m * 4 + 3]);
      }
      bar_target += gridDim.x;
      if (need_barrier) { if (grid_wait(P.bar, P.abort_flag, bar_target, &s_abort)) return; }
      else __syncthreads();
      GNAS_MARK(j + 1, 3);
    }
  }
  if (P.training && P.running && tid < 8) P.running[(tid >> 2) * H + n0 + (tid & 3)] = run[tid];
}

// ------------------------------------------------------------------ backward: per-step kernels
struct RhnBwd {
  int T, B, BP, H, t, node;           // node i+1 in 1..8 (elem) ; t current time step
  const float* ST; float* CHT; float* C0T; const float* HT; const float* h0T; const float* stat;
  const float* probs; const float* dout;
  float* dsT; float* dh_cur; float* dh_next;
  double* dpacc;
  RhnTable tb;
};

// dh[n][m] = dout[t][m][n]
static __global__ void __launch_bounds__(256) k_rhn_dout_T(const float* dout, float* dh, int B, int BP, int H, int t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H * BP) { const int n = i / BP, m = i - n * BP; dh[i] = m < B ? dout[((size_t)t * B + m) * H + n] : 0.f; }
}

// ds_k = dh/8 for k = 1..8, ds_0 = 0
static __global__ void __launch_bounds__(256) k_rhn_bwd_tinit(float* dsT, const float* dh, long long HB) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 9 * HB; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i / HB);
    dsT[i] = k == 0 ? 0.f : 0.125f * dh[i - (long long)k * HB];
  }
}

// node = i+1: dr = BN_bwd(ds_{i+1}); edge (i, j = blockIdx.y): dc, dh (in place over c, h), direct ds_j term,
// dprobs.  One warp per hidden column; the BN reduction is recomputed per edge (two loads per row).
static __global__ void __launch_bounds__(kThreads) k_rhn_bwd_elem(const __grid_constant__ RhnBwd P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 8 + warp;
  const int i = P.node - 1, j = blockIdx.y;
  const int B = P.B, BP = P.BP, H = P.H;
  const long long HB = (long long)H * BP;
  const int p = i * (i + 1) / 2 + j;
  __shared__ float red[8][4];
  float dpl[4] = {0.f, 0.f, 0.f, 0.f};
  if (n < H) {
    const size_t row = (size_t)n * BP;
    const float* dsr = P.dsT + (size_t)P.node * HB + row;
    const float* shr = P.ST + ((size_t)P.t * 9 + P.node) * HB + row;
    const float rstd = P.stat[((size_t)P.t * 9 + P.node) * 2 * H + H + n];
    const bool r0 = lane < B, r1 = lane + 32 < B;
    const float d0 = r0 ? dsr[lane] : 0.f, d1 = r1 ? dsr[lane + 32] : 0.f;
    const float h0 = r0 ? shr[lane] : 0.f, h1 = r1 ? shr[lane + 32] : 0.f;
    float* cp = P.CHT + (((size_t)P.t * 36 + p) * 2) * HB + row;
    float* hp = cp + HB;
    const float* sjr = P.ST + ((size_t)P.t * 9 + j) * HB + row;
    float* dsj = P.dsT + (size_t)j * HB + row;
    // every operand is fetched before the two batch reductions: one exposure to the load latency instead of two
    const float cv0 = r0 ? cp[lane] : 0.f, cv1 = r1 ? cp[lane + 32] : 0.f;
    const float hv0 = r0 ? hp[lane] : 0.f, hv1 = r1 ? hp[lane + 32] : 0.f;
    const float sj0 = r0 ? sjr[lane] : 0.f, sj1 = r1 ? sjr[lane + 32] : 0.f;
    float pk[4], q = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int ix = P.tb.pidx[p][k]; pk[k] = ix >= 0 ? P.probs[ix] : 0.f; q += pk[k]; }
    const float S1 = warp_sum(d0 + d1) / (float)B;
    const float S2 = warp_sum(d0 * h0 + d1 * h1) / (float)B;
    const float dr0 = rstd * (d0 - S1 - h0 * S2), dr1 = rstd * (d1 - S1 - h1 * S2);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = lane + 32 * half;
      if (m < BP) {
        float dc = 0.f, dh = 0.f;
        if (m < B) {
          const float dr = half ? dr1 : dr0;
          const float c = half ? cv1 : cv0, h = half ? hv1 : hv0, sj = half ? sj1 : sj0;
          const float g = sigmoidf_(c);
          float F = 0.f, Fp = 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (pk[k] != 0.f || P.tb.pidx[p][k] >= 0) {
              const float f = act_f(k, h);
              F = fmaf(pk[k], f, F); Fp = fmaf(pk[k], act_df(k, h), Fp);
              dpl[k] = fmaf(dr, sj + g * (f - sj), dpl[k]);
            }
          }
          dc = dr * g * (1.f - g) * (F - q * sj);
          dh = dr * g * Fp;
          atomicAdd(&dsj[m], dr * q * (1.f - g));     // the products of earlier nodes may still be landing in ds_j
        }
        cp[m] = dc; hp[m] = dh;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) { const float v = warp_sum(dpl[k]); if (lane == 0) red[warp][k] = v; }
  __syncthreads();
  if (threadIdx.x < 4) {
    const int k = threadIdx.x;
    if (P.tb.pidx[p][k] >= 0) {
      float v = 0.f;
      for (int w = 0; w < 8; ++w) v += red[w][k];
      atomicAdd(&P.dpacc[p * 4 + k], (double)v);
    }
  }
}

// stage 0: dr0 = BN_bwd(ds_0); dh_prev direct term (+ dout[t-1]); dc0, dh0 in place over C0T
static __global__ void __launch_bounds__(kThreads) k_rhn_bwd_elem0(const __grid_constant__ RhnBwd P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 8 + warp;
  const int B = P.B, BP = P.BP, H = P.H;
  const long long HB = (long long)H * BP;
  if (n >= H) return;
  const size_t row = (size_t)n * BP;
  const float* dsr = P.dsT + row;
  const float* shr = P.ST + ((size_t)P.t * 9) * HB + row;
  const float rstd = P.stat[((size_t)P.t * 9) * 2 * H + H + n];
  const bool r0 = lane < B, r1 = lane + 32 < B;
  const float d0 = r0 ? dsr[lane] : 0.f, d1 = r1 ? dsr[lane + 32] : 0.f;
  const float s0 = r0 ? shr[lane] : 0.f, s1 = r1 ? shr[lane + 32] : 0.f;
  const float S1 = warp_sum(d0 + d1) / (float)B;
  const float S2 = warp_sum(d0 * s0 + d1 * s1) / (float)B;
  float* cp = P.C0T + ((size_t)P.t * 2) * HB + row;
  float* hp = cp + HB;
  const float* hprev = (P.t > 0 ? P.HT + (size_t)(P.t - 1) * HB : P.h0T) + row;
  for (int half = 0; half < 2; ++half) {
    const int m = lane + 32 * half;
    if (m >= BP) continue;
    float dc = 0.f, dh = 0.f, dhp = 0.f;
    if (m < B) {
      const float dr = rstd * ((half ? d1 : d0) - S1 - (half ? s1 : s0) * S2);
      const float c = cp[m], h = hp[m], hv = hprev[m];
      const float g = sigmoidf_(c), v = tanhf(h);
      dhp = dr * (1.f - g);
      dc = dr * g * (1.f - g) * (v - hv);
      dh = dr * g * (1.f - v * v);
      if (P.t > 0) dhp += P.dout[((size_t)(P.t - 1) * B + m) * H + n];
    }
    cp[m] = dc; hp[m] = dh;
    P.dh_next[row + m] = dhp;
  }
}

// dst[c][r] = src[r][c]  (R x C -> C x R), batched over blockIdx.z
static __global__ void __launch_bounds__(256) k_transpose(const float* src, float* dst, int R, int C) {
  __shared__ float tile[32][33];
  src += (size_t)blockIdx.z * R * C; dst += (size_t)blockIdx.z * R * C;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) if (r0 + i < R && c0 + tx < C) tile[i][tx] = src[(size_t)(r0 + i) * C + c0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8) if (c0 + i < C && r0 + tx < R) dst[(size_t)(c0 + i) * R + r0 + tx] = tile[tx][i];
}

// out_z[n][m] += mask[n][m] * sum_k W[n][k] * D_z[k][m]   (split-K, atomics);  W is passed TRANSPOSED (WT[k][n])
// CTA tile: 64 rows n x BP columns m, K range = K / ksplit, cp.async double-buffered 32-row K chunks.
struct RhnGemm {
  int Nrows, K, BP, nz, ksplit, Hsplit;   // rows >= Hsplit go to out2/mask2 (stage 0: dx rows | dh rows)
  int z0;                                 // first problem of this launch: z = z0 + blockIdx.z
  const float* WT; int ldw;               // WT[k][n], row pitch ldw (= Nrows)
  const float* D[8]; float* out[8];
  const float* mask; float* out2; const float* mask2;
};
constexpr int kGemmKC = 32;
static __global__ void __launch_bounds__(kThreads) k_rhn_bwd_gemm(const __grid_constant__ RhnGemm P) {
  const int z = P.z0 + blockIdx.z;
  const int n0 = blockIdx.x * 64;
  const int kper = P.K / P.ksplit;
  const int kbeg = blockIdx.y * kper;
  const int BP = P.BP;
  const int tid = threadIdx.x;
  const int ncg = BP / 4;
  const int rq = tid / ncg, cg = tid - rq * ncg;
  const bool active = rq < 16;
  GNAS_DYN_SMEM(float, gsm);                                 // nch stages of { W [32][64] | D [32][BP] }
  const int stage_floats = kGemmKC * 64 + kGemmKC * BP;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const float* Dz = P.D[z];
  const int nch = kper / kGemmKC;
  const int ncols = min(64, P.Nrows - n0);           // multiple of 4 (H is a multiple of 32)
  auto issue = [&](int c) {
    const int k0 = kbeg + c * kGemmKC;
    float* wd = gsm + (size_t)c * stage_floats;
    float* dd = wd + kGemmKC * 64;
    for (int i = tid; i < kGemmKC * 16; i += kThreads) {
      const int k = i >> 4, q = i & 15;
      if (4 * q < ncols) cp_async16(wd + k * 64 + 4 * q, P.WT + (size_t)(k0 + k) * P.ldw + n0 + 4 * q);
    }
    for (int i = tid; i < kGemmKC * BP / 4; i += kThreads) cp_async16(dd + 4 * i, Dz + (size_t)k0 * BP + 4 * i);
    cp_async_commit();
  };
  // every chunk of this CTA's K range is requested up front (<= 4 x 16 KB): one L2 round trip, not one per chunk
  for (int c = 0; c < nch; ++c) issue(c);
  for (int c = 0; c < nch; ++c) {
    switch (nch - 1 - c) { case 0: cp_async_wait<0>(); break; case 1: cp_async_wait<1>(); break;
                           case 2: cp_async_wait<2>(); break; default: cp_async_wait<3>(); break; }
    __syncthreads();
    if (active) {
      const float* wr = gsm + (size_t)c * stage_floats + 4 * rq;
      const float* dr = gsm + (size_t)c * stage_floats + kGemmKC * 64 + 4 * cg;
      float4 w = *reinterpret_cast<const float4*>(wr);
      float4 d = *reinterpret_cast<const float4*>(dr);
#pragma unroll
      for (int k = 0; k < kGemmKC; ++k) {
        const float4 x = w, y = d;
        if (k + 1 < kGemmKC) {
          w = *reinterpret_cast<const float4*>(wr + (k + 1) * 64);
          d = *reinterpret_cast<const float4*>(dr + (k + 1) * BP);
        }
        acc[0][0] = fmaf(x.x, y.x, acc[0][0]); acc[0][1] = fmaf(x.x, y.y, acc[0][1]); acc[0][2] = fmaf(x.x, y.z, acc[0][2]); acc[0][3] = fmaf(x.x, y.w, acc[0][3]);
        acc[1][0] = fmaf(x.y, y.x, acc[1][0]); acc[1][1] = fmaf(x.y, y.y, acc[1][1]); acc[1][2] = fmaf(x.y, y.z, acc[1][2]); acc[1][3] = fmaf(x.y, y.w, acc[1][3]);
        acc[2][0] = fmaf(x.z, y.x, acc[2][0]); acc[2][1] = fmaf(x.z, y.y, acc[2][1]); acc[2][2] = fmaf(x.z, y.z, acc[2][2]); acc[2][3] = fmaf(x.z, y.w, acc[2][3]);
        acc[3][0] = fmaf(x.w, y.x, acc[3][0]); acc[3][1] = fmaf(x.w, y.y, acc[3][1]); acc[3][2] = fmaf(x.w, y.z, acc[3][2]); acc[3][3] = fmaf(x.w, y.w, acc[3][3]);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int n = n0 + 4 * rq + a;
      if (n >= P.Nrows) continue;
      float* o; const float* mk;
      if (n < P.Hsplit) { o = P.out[z] + (size_t)n * BP; mk = P.mask + (size_t)n * BP; }
      else { o = P.out2 + (size_t)(n - P.Hsplit) * BP; mk = P.mask2 + (size_t)(n - P.Hsplit) * BP; }
      const float4 mv = *reinterpret_cast<const float4*>(mk + 4 * cg);
      red_add_v4(o + 4 * cg, mv.x * acc[a][0], mv.y * acc[a][1], mv.z * acc[a][2], mv.w * acc[a][3]);
    }
  }
}

// dW[n][k'] += sum over blocks, sum_m A[n][m] * D[k'][m];  A = mask .* S (or the stage-0 input rows)
struct RhnWg {
  int T, B, BP, H, tsplit, is_w0;
  const float* ST; const float* CHT; const float* C0T; const float* HT; const float* h0T;
  const float* xmT; const float* hmT;
  float* dW0; float* dWs;
};
static __global__ void __launch_bounds__(kThreads) k_rhn_wgrad(const __grid_constant__ RhnWg P) {
  const int H = P.H, BP = P.BP, B = P.B;
  const long long HB = (long long)H * BP;
  const int n0 = blockIdx.x * 64, k0 = blockIdx.y * 64;
  const int iw = P.is_w0 ? 0 : (int)(blockIdx.z % 8);
  const int ts = P.is_w0 ? blockIdx.z : blockIdx.z / 8;
  const int tper = (P.T + P.tsplit - 1) / P.tsplit;
  const int tb = ts * tper, te = min(P.T, tb + tper);
  const int Nrows = P.is_w0 ? 2 * H : H;
  const int tid = threadIdx.x, tr = tid >> 4, tc = tid & 15;
  __shared__ float As[64][68];   // [m][n]
  __shared__ float Ds[64][68];   // [m][k']
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const int nblk = P.is_w0 ? 1 : iw + 1;
  for (int t = tb; t < te; ++t) {
    for (int j = 0; j < nblk; ++j) {
      const float* Dsrc = P.is_w0 ? P.C0T + ((size_t)t * 2) * HB : P.CHT + (((size_t)t * 36 + iw * (iw + 1) / 2 + j) * 2) * HB;
      for (int i = tid; i < 64 * BP; i += kThreads) {
        const int r = i / BP, m = i - r * BP;
        const int n = n0 + r;
        float a = 0.f;
        if (n < Nrows && m < B) {
          if (P.is_w0) {
            if (n < H) a = P.xmT[(size_t)t * HB + (size_t)n * BP + m];
            else {
              const size_t o = (size_t)(n - H) * BP + m;
              a = (t > 0 ? P.HT[(size_t)(t - 1) * HB + o] : P.h0T[o]) * P.hmT[o];
            }
          } else {
            const size_t o = (size_t)n * BP + m;
            a = P.ST[((size_t)t * 9 + j) * HB + o] * P.hmT[o];
          }
        }
        As[m][r] = a;
        Ds[m][r] = (m < B) ? Dsrc[(size_t)(k0 + r) * BP + m] : 0.f;
      }
      __syncthreads();
      for (int m = 0; m < BP; ++m) {
        const float4 a = *reinterpret_cast<const float4*>(&As[m][4 * tr]);
        const float4 d = *reinterpret_cast<const float4*>(&Ds[m][4 * tc]);
        acc[0][0] = fmaf(a.x, d.x, acc[0][0]); acc[0][1] = fmaf(a.x, d.y, acc[0][1]); acc[0][2] = fmaf(a.x, d.z, acc[0][2]); acc[0][3] = fmaf(a.x, d.w, acc[0][3]);
        acc[1][0] = fmaf(a.y, d.x, acc[1][0]); acc[1][1] = fmaf(a.y, d.y, acc[1][1]); acc[1][2] = fmaf(a.y, d.z, acc[1][2]); acc[1][3] = fmaf(a.y, d.w, acc[1][3]);
        acc[2][0] = fmaf(a.z, d.x, acc[2][0]); acc[2][1] = fmaf(a.z, d.y, acc[2][1]); acc[2][2] = fmaf(a.z, d.z, acc[2][2]); acc[2][3] = fmaf(a.z, d.w, acc[2][3]);
        acc[3][0] = fmaf(a.w, d.x, acc[3][0]); acc[3][1] = fmaf(a.w, d.y, acc[3][1]); acc[3][2] = fmaf(a.w, d.z, acc[3][2]); acc[3][3] = fmaf(a.w, d.w, acc[3][3]);
      }
      __syncthreads();
    }
  }
  float* dW = P.is_w0 ? P.dW0 : P.dWs + (size_t)iw * H * 2 * H;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int n = n0 + 4 * tr + a;
    if (n >= Nrows) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int k = k0 + 4 * tc + b;
      if (k < 2 * H) atomicAdd(&dW[(size_t)n * 2 * H + k], acc[a][b]);
    }
  }
}

static __global__ void k_rhn_dprobs(const double* dpacc, float* dprobs, int nprobs, RhnTable tb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nprobs) dprobs[i] = 0.f;
  __syncthreads();
  if (i < 36 * 4) {
    const int ix = tb.pidx[i >> 2][i & 3];
    if (ix >= 0) dprobs[ix] = (float)dpacc[i];
  }
}

static int rhn_validate(const GnasRhnDesc* d) {
  if (!d) { set_error("null descriptor"); return GNAS_ERR_ARG; }
  if (d->T <= 0 || d->B <= 0 || d->H <= 0) { set_error("rhn: bad shape"); return GNAS_ERR_SHAPE; }
  if (d->H % kKCH != 0) { set_error("rhn: hidden size must be a multiple of 32"); return GNAS_ERR_SHAPE; }
  if ((d->B + 3) / 4 * 4 > kMaxBP) { set_error("rhn: at most 64 batch rows per GPU"); return GNAS_ERR_SHAPE; }
  return 0;
}

static inline int ewb(long long total) {
  long long nb = (total + 255) / 256;
  return (int)(nb < 1 ? 1 : (nb > 148 * 8 ? 148 * 8 : nb));
}

}  // namespace gnas

using namespace gnas;

extern "C" int gnas_rhn_workspace(const GnasRhnDesc* d, int64_t* fwd_floats, int64_t* bwd_floats) {
  int rc = rhn_validate(d);
  if (rc) return rc;
  RhnPlan p = rhn_plan(d);
  if (fwd_floats) *fwd_floats = p.fwd_total;
  if (bwd_floats) *bwd_floats = p.bwd_total;
  return 0;
}

extern "C" int gnas_rhn_forward(const GnasRhnDesc* d, const float* x, const float* h0, const float* W0,
                                const float* Ws, const float* probs, const float* x_mask, const float* h_mask,
                                float* bn_running, float* out, float* ws, void* stream) {
  int rc = rhn_validate(d);
  if (rc) return rc;
  if (!x || !h0 || !W0 || !Ws || !probs || !out || !ws) { set_error("rhn_forward: null pointer"); return GNAS_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  RhnPlan p = rhn_plan(d);
  const int H = d->H, BP = p.BP;
  const long long HB = (long long)H * BP;
  RhnPro pro;
  pro.T = d->T; pro.B = d->B; pro.BP = BP; pro.H = H;
  pro.x = x; pro.h0 = h0; pro.xm = d->training ? x_mask : nullptr; pro.hm = d->training ? h_mask : nullptr;
  pro.xmT = ws + p.xmT; pro.xkT = ws + p.xkT; pro.hmT = ws + p.hmT; pro.h0T = ws + p.h0T; pro.ring9 = ws + p.ring + 9 * HB;
  GNAS_LAUNCH(k_rhn_prologue, dim3(ewb((long long)(d->T + 3) * HB)), dim3(256), 0, st, pro);
  cudaMemsetAsync(ws + p.bar, 0, 256 * sizeof(float), st);
  RhnFwd f;
  f.T = d->T; f.B = d->B; f.BP = BP; f.H = H; f.training = d->training; f.momentum = d->bn_momentum;
  f.xmT = ws + p.xmT; f.hmT = ws + p.hmT; f.h0T = ws + p.h0T; f.W0 = W0; f.Ws = Ws; f.probs = probs;
  f.ring = ws + p.ring; f.ST = ws + p.ST; f.CHT = ws + p.CHT; f.C0T = ws + p.C0T; f.HT = ws + p.HT; f.stat = ws + p.stat;
  f.out = out; f.running = bn_running;
  f.bar = reinterpret_cast<unsigned*>(ws + p.bar); f.abort_flag = reinterpret_cast<int*>(ws + p.bar) + 8;
  rhn_table(d, f.tb);
  {
    int nsm = 1 << 30;
#ifndef GNAS_EMUL
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
#endif
    if (rhn_tc_wanted() && rtc::supported(d->B, H, nsm)) {
      const rtc::Plan tp = rtc::plan(d->T, H);
      float* tcw = ws + p.tc;
      rtc::Pre pre;
      pre.T = d->T; pre.H = H; pre.BP = BP; pre.xmT = ws + p.xmT; pre.h0T = ws + p.h0T; pre.hmT = ws + p.hmT;
      pre.xc = tcw + tp.xc; pre.ring9 = tcw + tp.ringc + 9LL * tp.nch * rtc::S_CHUNK;
      GNAS_LAUNCH(rtc::k_rhn_tc_pre, dim3(ewb((long long)(d->T + 1) * H * rtc::BPT)), dim3(256), 0, st, pre);
      rtc::Pack pk;
      pk.H = H; pk.W0 = W0; pk.Ws = Ws; pk.wp = tcw + tp.wp; pk.wp_cta = tp.wp_cta;
      GNAS_LAUNCH(rtc::k_rhn_tc_pack, dim3(H / rtc::NC), dim3(256), 0, st, pk);
      rtc::Fwd f;
      f.T = d->T; f.B = d->B; f.BP = BP; f.H = H; f.training = d->training; f.momentum = d->bn_momentum;
      f.hmT = ws + p.hmT; f.h0T = ws + p.h0T; f.probs = probs;
      f.xc = tcw + tp.xc; f.ringc = tcw + tp.ringc; f.wp = tcw + tp.wp; f.wp_cta = tp.wp_cta;
      f.ST = ws + p.ST; f.CHT = ws + p.CHT; f.C0T = ws + p.C0T; f.HT = ws + p.HT; f.stat = ws + p.stat;
      f.out = out; f.running = bn_running; f.bar = reinterpret_cast<unsigned*>(ws + p.bar);
      rhn_table(d, f.tb);
      const size_t tsm = rtc::smem_bytes();
      const int tgrid = H / rtc::NC;
#ifdef GNAS_EMUL
      GNAS_LAUNCH_COOP(rtc::k_rhn_fwd_tc, dim3(tgrid), dim3(rtc::THREADS), tsm, st, f);
#else
      static int prepared[kMaxDevices] = {0};
      if (need_prepare(prepared, (int)tsm)) {
        cudaError_t e = cudaFuncSetAttribute(rtc::k_rhn_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
        if (e != cudaSuccess) { set_error("rhn_forward: shared memory request rejected (tensor-core kernel)"); return (int)e; }
      }
      void* targs[] = {(void*)&f};
      prof_before("k_rhn_fwd_tc", "", st);
      cudaError_t e = cudaLaunchCooperativeKernel((void*)rtc::k_rhn_fwd_tc, dim3(tgrid), dim3(rtc::THREADS), targs, tsm, st);
      prof_after(st);
      if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
#endif
      return check_launch("gnas_rhn_forward");
    }
  }
  const size_t smem = sizeof(float) * ((size_t)2 * H * 8 + (size_t)8 * H * 8 + (size_t)kStages * kKCH * BP +
                                       8 * BP * 4 + 9 * BP * 4 + 3 * BP * 4 + 36 * 5 + 8 + 8);
  const int grid = H / kNC;
#ifdef GNAS_EMUL
  GNAS_LAUNCH_COOP(k_rhn_forward, dim3(grid), dim3(kRhnThreads), smem, st, f);
#else
  {
    static int prepared[kMaxDevices] = {0};
    if (need_prepare(prepared, (int)smem)) {
      cudaError_t e = cudaFuncSetAttribute(k_rhn_forward, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_error("rhn_forward: shared memory request rejected"); return (int)e; }
    }
    int dev = 0, nsm = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rhn_forward, kRhnThreads, smem);
    if (per_sm * nsm < grid) { set_error("rhn_forward: grid does not fit co-resident (hidden size too large)"); return GNAS_ERR_UNSUPPORTED; }
    void* args[] = {(void*)&f};
    prof_before("k_rhn_forward", "", st);
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_rhn_forward, dim3(grid), dim3(kRhnThreads), args, smem, st);
    prof_after(st);
    if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); return (int)e; }
  }
#endif
  return check_launch("gnas_rhn_forward");
}

extern "C" int gnas_rhn_backward(const GnasRhnDesc* d, const float* x, const float* h0, const float* W0,
                                 const float* Ws, const float* probs, const float* x_mask, const float* h_mask,
                                 const float* dout, float* ws, float* scratch, float* dx, float* dh0,
                                 float* dW0, float* dWs, float* dprobs, int need_wgrad, void* stream) {
  int rc = rhn_validate(d);
  if (rc) return rc;
  (void)x; (void)h0; (void)x_mask; (void)h_mask;
  if (!W0 || !Ws || !probs || !dout || !ws || !scratch || !dx || !dh0 || !dprobs || (need_wgrad && (!dW0 || !dWs))) {
    set_error("rhn_backward: null pointer"); return GNAS_ERR_ARG;
  }
  if (!d->training) {   // the kernels differentiate batch-statistics BatchNorm only
    set_error("rhn_backward: the forward ran in eval mode (running statistics); its backward is not implemented");
    return GNAS_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  RhnPlan p = rhn_plan(d);
  const int H = d->H, BP = p.BP, B = d->B, T = d->T;
  const long long HB = (long long)H * BP;
  RhnTable tb;
  rhn_table(d, tb);
  cudaMemsetAsync(scratch, 0, sizeof(float) * (size_t)p.bwd_zero, st);
  GNAS_LAUNCH(k_transpose, dim3((2 * H + 31) / 32, (H + 31) / 32, 8), dim3(256), 0, st, Ws, scratch + p.WsT, H, 2 * H);
  GNAS_LAUNCH(k_transpose, dim3((2 * H + 31) / 32, (2 * H + 31) / 32, 1), dim3(256), 0, st, W0, scratch + p.W0T, 2 * H, 2 * H);
  float* dhbuf[2] = {scratch + p.dhT, scratch + p.dhT + HB};
  // dh for the last step = dout[T-1]^T
  GNAS_LAUNCH(k_rhn_dout_T, dim3((H * BP + 255) / 256), dim3(256), 0, st, dout, dhbuf[0], B, BP, H, T - 1);
  RhnBwd bp;
  bp.T = T; bp.B = B; bp.BP = BP; bp.H = H;
  bp.ST = ws + p.ST; bp.CHT = ws + p.CHT; bp.C0T = ws + p.C0T; bp.HT = ws + p.HT; bp.h0T = ws + p.h0T; bp.stat = ws + p.stat;
  bp.probs = probs; bp.dout = dout; bp.dsT = scratch + p.dsT; bp.dpacc = reinterpret_cast<double*>(scratch + p.dpacc);
  bp.tb = tb;
  const int eb = (H + 7) / 8;
  int ksplit = 1;                       // split K = 2H into CTAs of at most 4 chunks of 32 (all prefetched at once)
  static int kchunks = -1;
  if (kchunks < 0) { const char* e = getenv("GNAS_RHN_KCH"); kchunks = e ? atoi(e) : 2; }   // 2 x 32 rows of K per CTA: 1024 CTAs for node 7, one round at 7 CTAs per SM (4: 512 CTAs, two rounds at 3 per SM)
  while ((2 * H) / ksplit > kchunks * kGemmKC && (2 * H) % (ksplit * 2 * kGemmKC) == 0) ksplit *= 2;
  const int nch_g = (2 * H) / ksplit / kGemmKC;
  const size_t gsm = sizeof(float) * (size_t)nch_g * (kGemmKC * 64 + kGemmKC * BP);
#ifndef GNAS_EMUL
  { static int prepared[kMaxDevices] = {0};
    if (need_prepare(prepared, (int)gsm)) cudaFuncSetAttribute(k_rhn_bwd_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm); }
#endif
  if (nch_g > 4 || nch_g < 1) { set_error("rhn_backward: unsupported hidden size for the GEMM split"); return GNAS_ERR_UNSUPPORTED; }
  const SideStream side = side_stream(0);
  for (int t = T - 1; t >= 0; --t) {
    float* dh_cur = dhbuf[(T - 1 - t) & 1];
    float* dh_next = dhbuf[(T - t) & 1];
    bp.t = t; bp.dh_cur = dh_cur; bp.dh_next = dh_next;
    GNAS_LAUNCH(k_rhn_bwd_tinit, dim3(ewb(9 * HB)), dim3(256), 0, st, scratch + p.dsT, (const float*)dh_cur, HB);
    for (int i = 7; i >= 0; --i) {
      bp.node = i + 1;
      GNAS_LAUNCH(k_rhn_bwd_elem, dim3(eb, i + 1), dim3(kThreads), 0, st, bp);
      RhnGemm g;
      g.Nrows = H; g.K = 2 * H; g.BP = BP; g.nz = i + 1; g.ksplit = ksplit; g.Hsplit = H;
      g.WT = scratch + p.WsT + (size_t)i * H * 2 * H; g.ldw = H;
      for (int j = 0; j <= i; ++j) {
        g.D[j] = ws + p.CHT + (((size_t)t * 36 + i * (i + 1) / 2 + j) * 2) * HB;
        g.out[j] = scratch + p.dsT + (size_t)j * HB;
      }
      g.mask = ws + p.hmT; g.out2 = nullptr; g.mask2 = nullptr;
      // Only the product of edge (i, i) feeds the next stage (ds_i is complete once it has landed); the products
      // of the edges (i, j < i) are needed at stage j at the earliest, so they run on a side stream next to the
      // following elementwise kernel and critical product instead of in front of them.
      if (side.st) {
        g.z0 = i; g.nz = 1;
        if (i > 0) { cudaEventRecord(side.fork, st); cudaStreamWaitEvent(side.st, side.fork, 0); }
        GNAS_LAUNCH(k_rhn_bwd_gemm, dim3((H + 63) / 64, g.ksplit, 1), dim3(kThreads), gsm, st, g);
        // stage i - 1 reads ds_i: every side product into it was issued at stage i + 1 or before
        if (i < 7) cudaStreamWaitEvent(st, side.join[(i + 1) & 1], 0);
        if (i > 0) {
          g.z0 = 0; g.nz = i;
          GNAS_LAUNCH(k_rhn_bwd_gemm, dim3((H + 63) / 64, g.ksplit, i), dim3(kThreads), gsm, side.st, g);
          cudaEventRecord(side.join[i & 1], side.st);
        }
      } else {
        g.z0 = 0;
        GNAS_LAUNCH(k_rhn_bwd_gemm, dim3((H + 63) / 64, g.ksplit, i + 1), dim3(kThreads), gsm, st, g);
      }
    }
    GNAS_LAUNCH(k_rhn_bwd_elem0, dim3(eb), dim3(kThreads), 0, st, bp);
    RhnGemm g;
    g.Nrows = 2 * H; g.K = 2 * H; g.BP = BP; g.nz = 1; g.ksplit = ksplit; g.Hsplit = H;
    g.WT = scratch + p.W0T; g.ldw = 2 * H;
    g.D[0] = ws + p.C0T + ((size_t)t * 2) * HB;
    g.out[0] = scratch + p.dxT + (size_t)t * HB; g.mask = ws + p.xkT;
    g.out2 = dh_next; g.mask2 = ws + p.hmT; g.z0 = 0;
    GNAS_LAUNCH(k_rhn_bwd_gemm, dim3((2 * H + 63) / 64, g.ksplit, 1), dim3(kThreads), gsm, st, g);
  }
  // outputs: dx [T][B][H], dh0 [B][H], dprobs
  GNAS_LAUNCH(k_rhn_untranspose, dim3(ewb((long long)T * B * H)), dim3(256), 0, st, (const float*)(scratch + p.dxT), dx, T, B, BP, H);
  GNAS_LAUNCH(k_rhn_untranspose, dim3(ewb((long long)B * H)), dim3(256), 0, st, (const float*)dhbuf[T & 1], dh0, 1, B, BP, H);
  GNAS_LAUNCH(k_rhn_dprobs, dim3(1), dim3(256), 0, st, (const double*)(scratch + p.dpacc), dprobs, 36 * d->n_prims, tb);
#ifndef GNAS_EMUL
  if (need_wgrad && tc_enabled() && tc::rhn_wgrad_tc_supported(B, BP, H)) {
    tc::RhnWgTc w;
    w.T = T; w.B = B; w.BP = BP; w.H = H;
    w.ST = ws + p.ST; w.CHT = ws + p.CHT; w.C0T = ws + p.C0T; w.HT = ws + p.HT; w.h0T = ws + p.h0T;
    w.xmT = ws + p.xmT; w.hmT = ws + p.hmT; w.dW0 = dW0; w.dWs = dWs;
    tc::launch_rhn_wgrad_tc(w, st);
  } else
#endif
  if (need_wgrad) {
    RhnWg w;
    w.T = T; w.B = B; w.BP = BP; w.H = H;
    w.ST = ws + p.ST; w.CHT = ws + p.CHT; w.C0T = ws + p.C0T; w.HT = ws + p.HT; w.h0T = ws + p.h0T;
    w.xmT = ws + p.xmT; w.hmT = ws + p.hmT; w.dW0 = dW0; w.dWs = dWs;
    w.tsplit = T >= 6 ? 6 : 1;
    w.is_w0 = 0;
    GNAS_LAUNCH(k_rhn_wgrad, dim3((H + 63) / 64, (2 * H + 63) / 64, 8 * w.tsplit), dim3(kThreads), 0, st, w);
    w.is_w0 = 1;
    GNAS_LAUNCH(k_rhn_wgrad, dim3((2 * H + 63) / 64, (2 * H + 63) / 64, w.tsplit), dim3(kThreads), 0, st, w);
  }
  return check_launch("gnas_rhn_backward");
}

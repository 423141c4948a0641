// Forward kernels of the MixedOp / CNN cell path (fp32, NCL layout).
//
// Reference arithmetic being replaced (all executed there by ATen/cuDNN calls):
//   generalNAS_tools/operations_14_9.py:17-30,67-81,124-135,164-206,280-300
//   darts_tools/model_discCNN.py:46-93 (MixedOp), :96-201 (CNN_Cell_search)
// Every candidate ends in training-mode BatchNorm1d(affine=False), so an edge is
// evaluated in phases: raw convolution outputs + per-channel (sum, sumsq) ->
// finalize (mean, rstd) -> consumers normalise on load (stage 2) or in the mix.
#pragma once
#include "common.cuh"

namespace gnas {

// ------------------------------------------------------------------ weight packing
struct PackJob { const float* src; float* dst; int Cout, Cin, KW, mode; };  // mode 0: [ci][t][co], 1: [co][t][ci]
constexpr int kMaxPack = 160;
struct PackBatch { int n; PackJob j[kMaxPack]; };

static __global__ void __launch_bounds__(256) k_pack_weights(const __grid_constant__ PackBatch P) {
  const PackJob& J = P.j[blockIdx.y];
  const int total = J.Cout * J.Cin * J.KW;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i % J.KW;
    const int ci = (i / J.KW) % J.Cin;
    const int co = i / (J.KW * J.Cin);
    const int o = J.mode == 0 ? (ci * J.KW + t) * J.Cout + co : (co * J.KW + t) * J.Cin + ci;
    J.dst[o] = J.src[i];
  }
}

// ------------------------------------------------------------------ dense KW-tap conv
struct DenseJob {
  const float* x; long long xbs;   // input [B][Cin][Lin], batch stride in floats
  const float* in_stat;            // mean[Cin] | rstd[Cin] (T_BN_RELU / T_BN) or null
  const float* wp;                 // packed [Cin][KW][Cout]
  float* z; long long zbs;         // raw output [B][Cout][Lout]
  double* acc;                     // sum[Cout] | sumsq[Cout], or null
  int Cin, Lin, Lout, tin;
};
constexpr int kMaxDense = 32;
struct DenseBatch { int n, B, Cout, TL, pad, KC; DenseJob j[kMaxDense]; };

template <int TM> struct WVec;
template <> struct WVec<2> { static __device__ __forceinline__ void ld(const float* p, float* w) { float2 v = *reinterpret_cast<const float2*>(p); w[0] = v.x; w[1] = v.y; } };
template <> struct WVec<4> { static __device__ __forceinline__ void ld(const float* p, float* w) { float4 v = *reinterpret_cast<const float4*>(p); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; } };
template <> struct WVec<8> { static __device__ __forceinline__ void ld(const float* p, float* w) { WVec<4>::ld(p, w); WVec<4>::ld(p + 4, w + 4); } };

// z[b,co,l] = sum_{ci,t} W[co,ci,t] * T(x)[b,ci,l*S+t-pad]  (+ per-channel sum / sumsq of z)
// CTA: one batch row, TL output positions, all Cout channels; thread tile TM x 4.
template <int KW, int S, int TM>
__global__ void __launch_bounds__(kThreads) k_dense_fwd(const __grid_constant__ DenseBatch P) {
  const DenseJob& J = P.j[blockIdx.z];
  const int TL = P.TL, l0 = blockIdx.x * TL;
  if (l0 >= J.Lout) return;
  const int b = blockIdx.y, Cout = P.Cout, KC = P.KC;
  const int ncog = Cout / TM, nlg = TL / 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cog = tid / nlg, lg = tid - cog * nlg;
  const bool active = cog < ncog;
  constexpr int NX = 3 * S + KW;
  constexpr int NX4 = (NX + 3) / 4;
  const int Wu = (TL - 1) * S + KW;             // used width of an xs row
  const int Wp = (Wu + 3) / 4 * 4 + 4;          // padded pitch (window over-read stays inside)
  // Two-stage software pipeline over chunks of KC input channels: the weights of chunk c+1 arrive by
  // cp.async and the raw input rows of chunk c+1 are fetched into registers while chunk c is multiplied;
  // the input transform (ReLU / BN+ReLU) is applied when the registers are stored to shared memory.
  constexpr int NPRE = 10;                      // input elements a lane may hold in flight per row
  GNAS_DYN_SMEM(float, smem);
  const int xs_sz = KC * Wp, ws_sz = KC * KW * Cout;
  float* xsb = smem;                            // [2][KC][Wp]
  float* wsb = xsb + 2 * xs_sz;                 // [2][KC][KW][Cout]
  // per-CTA statistics are summed in double from the per-warp partials on: the float result no longer depends
  // on the order in which the warps arrive (deterministic BN statistics, hence reproducible ReLU decisions)
  double* ssum = reinterpret_cast<double*>(wsb + 2 * ws_sz);   // [2][Cout]
  for (int i = tid; i < 2 * Cout; i += kThreads) ssum[i] = 0.0;
  float acc[TM][4];
#pragma unroll
  for (int m = 0; m < TM; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[m][n] = 0.f;
  const int in0 = l0 * S - P.pad;
  const float* xb = J.x + (long long)b * J.xbs;
  const int nchunks = J.Cin / KC;
  const bool prefetch = KC <= kThreads / 32 && Wp <= 32 * NPRE;     // one row per warp, <= NPRE elements per lane
  auto load_w = [&](int c, int buf) {
    const float* src = J.wp + (long long)c * ws_sz;
    float* dst = wsb + buf * ws_sz;
    for (int i = 4 * tid; i < ws_sz; i += 4 * kThreads) gnas_cp_async16(dst + i, src + i);
    gnas_cp_async_commit();
  };
  auto load_x_sync = [&](int c, int buf) {
    float* xs = xsb + buf * xs_sz;
    for (int ci = warp; ci < KC; ci += kThreads / 32) {
      const float* xr = xb + (long long)(c * KC + ci) * J.Lin;
      float mean = 0.f, rstd = 1.f;
      if (J.tin >= T_BN_RELU) { mean = J.in_stat[c * KC + ci]; rstd = J.in_stat[J.Cin + c * KC + ci]; }
      _Pragma("unroll 4")
      for (int p = lane; p < Wp; p += 32) {
        const int pos = in0 + p;
        float v = 0.f;
        if (p < Wu && pos >= 0 && pos < J.Lin) v = apply_tin(xr[pos], J.tin, mean, rstd);
        xs[ci * Wp + p] = v;
      }
    }
  };
  float xpre[NPRE];
  auto fetch_x = [&](int c) {
    if (warp < KC) {
      const float* xr = xb + (long long)(c * KC + warp) * J.Lin;
#pragma unroll
      for (int q = 0; q < NPRE; ++q) {
        const int p = lane + 32 * q, pos = in0 + p;
        xpre[q] = (p < Wu && pos >= 0 && pos < J.Lin) ? xr[pos] : 0.f;
      }
    }
  };
  auto store_x = [&](int c, int buf) {
    if (warp < KC) {
      float* xs = xsb + buf * xs_sz + warp * Wp;
      float mean = 0.f, rstd = 1.f;
      if (J.tin >= T_BN_RELU) { mean = J.in_stat[c * KC + warp]; rstd = J.in_stat[J.Cin + c * KC + warp]; }
#pragma unroll
      for (int q = 0; q < NPRE; ++q) {
        const int p = lane + 32 * q, pos = in0 + p;
        if (p < Wp) xs[p] = (p < Wu && pos >= 0 && pos < J.Lin) ? apply_tin(xpre[q], J.tin, mean, rstd) : 0.f;
      }
    }
  };
  load_w(0, 0);
  load_x_sync(0, 0);
  gnas_cp_async_wait_all();
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const int cur = c & 1, nxt = cur ^ 1;
    const bool more = c + 1 < nchunks;
    if (more) { load_w(c + 1, nxt); if (prefetch) fetch_x(c + 1); }
    if (active) {
      const float* xs = xsb + cur * xs_sz;
      const float* ws = wsb + cur * ws_sz;
      for (int ci = 0; ci < KC; ++ci) {
        float xr[NX4 * 4];
        const float4* xp = reinterpret_cast<const float4*>(xs + ci * Wp + lg * 4 * S);
#pragma unroll
        for (int q = 0; q < NX4; ++q) { float4 v = xp[q]; xr[4 * q] = v.x; xr[4 * q + 1] = v.y; xr[4 * q + 2] = v.z; xr[4 * q + 3] = v.w; }
        const float* wrow = ws + ci * KW * Cout + cog * TM;
#pragma unroll
        for (int t = 0; t < KW; ++t) {
          float w[TM];
          WVec<TM>::ld(wrow + t * Cout, w);
#pragma unroll
          for (int m = 0; m < TM; ++m)
#pragma unroll
            for (int n = 0; n < 4; ++n) acc[m][n] = fmaf(w[m], xr[n * S + t], acc[m][n]);
        }
      }
    }
    if (more) {
      if (prefetch) store_x(c + 1, nxt); else load_x_sync(c + 1, nxt);
      gnas_cp_async_wait_all();
    }
    __syncthreads();
  }
  // ---- epilogue: store raw z, accumulate statistics
  const int lbase = l0 + 4 * lg;
  float* zb = J.z + (long long)b * J.zbs;
  const bool vec = active && (J.Lout % 4 == 0) && (J.zbs % 4 == 0) && (lbase + 3 < J.Lout);
  const bool pow2 = (nlg & (nlg - 1)) == 0;
  const int width = nlg < 32 ? nlg : 32;
#pragma unroll
  for (int m = 0; m < TM; ++m) {
    const int co = cog * TM + m;
    float s = 0.f, q = 0.f;
    if (active) {
      float* zr = zb + (long long)co * J.Lout;
      if (vec) {
        *reinterpret_cast<float4*>(zr + lbase) = make_float4(acc[m][0], acc[m][1], acc[m][2], acc[m][3]);
#pragma unroll
        for (int n = 0; n < 4; ++n) { s += acc[m][n]; q += acc[m][n] * acc[m][n]; }
      } else {
#pragma unroll
        for (int n = 0; n < 4; ++n)
          if (lbase + n < J.Lout) { zr[lbase + n] = acc[m][n]; s += acc[m][n]; q += acc[m][n] * acc[m][n]; }
      }
    }
    if (J.acc != nullptr) {
      if (pow2) {
        for (int o = width >> 1; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if (active && (lg & (width - 1)) == 0) { atomicAdd(&ssum[co], (double)s); atomicAdd(&ssum[Cout + co], (double)q); }
      } else if (active) {
        atomicAdd(&ssum[co], (double)s); atomicAdd(&ssum[Cout + co], (double)q);
      }
    }
  }
  if (J.acc != nullptr) {
    __syncthreads();
    for (int i = tid; i < 2 * Cout; i += kThreads) atomicAdd(&J.acc[i], ssum[i]);
  }
}

// ------------------------------------------------------------------ depthwise conv
struct DwJob {
  const float* x; long long xbs;
  const float* in_stat;            // mean | rstd of the producer BN (T_BN_RELU) or null
  const float* wd;                 // [C][K]
  float* u; long long ubs;         // output [B][C][Lout]
  int Lin, Lout, tin;
  int K;                           // 15 or 9 (k_dw2_fwd picks the instance per job)
};
constexpr int kMaxDw = 16;
struct DwBatch { int n, B, C, pad, nseg; DwJob j[kMaxDw]; };

// A row of outputs is cut into segments of SEGO positions (whole rows left the C = 8 cells with 1.5 CTAs per SM)
template <int K, int D, int S>
struct DwFwdGeom {
  static constexpr int SEGO = S == 1 ? 256 : (S == 2 ? 128 : 84);              // multiples of 4
  static constexpr int WP = (SEGO * S + (K - 1) * D + 3 * S + 8 + 3) / 4 * 4;   // staged inputs: index p <-> pos = l0*S - pad + p
};

// u[b,c,l] = sum_t wd[c,t] * T(x)[b,c,l*S+t*D-pad];  CTA: batch row b, 8 channels (one per warp), one output segment
template <int K, int D, int S>
__device__ __forceinline__ void dw_fwd_body(const DwBatch& P) {
  using G = DwFwdGeom<K, D, S>;
  constexpr int SEGO = G::SEGO, WP = G::WP, PADC = (K - 1) * D / 2;
  const DwJob& J = P.j[blockIdx.z];
  const int b = blockIdx.y / P.nseg, seg = blockIdx.y - b * P.nseg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 8 + warp;
  const int l0 = seg * SEGO, l1 = min(J.Lout, l0 + SEGO);
  constexpr int NX = 3 * S + (K - 1) * D + 1;
  constexpr int NX4 = (NX + 3) / 4;
  __shared__ __align__(16) float rows[8][WP];
  __shared__ __align__(16) float outs[8][SEGO];
  float* row = rows[warp];
  if (c >= P.C || l0 >= J.Lout) return;
  {
    const float* xr = J.x + (long long)b * J.xbs + (long long)c * J.Lin;
    float mean = 0.f, rstd = 1.f;
    if (J.tin >= T_BN_RELU) { mean = J.in_stat[c]; rstd = J.in_stat[P.C + c]; }
    float xv[(WP + 31) / 32];
#pragma unroll
    for (int q = 0; q < (WP + 31) / 32; ++q) {
      const int pos = l0 * S - PADC + lane + 32 * q;
      xv[q] = (pos >= 0 && pos < J.Lin) ? xr[pos] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < (WP + 31) / 32; ++q) {
      const int p = lane + 32 * q, pos = l0 * S - PADC + p;
      if (p < WP) row[p] = (pos >= 0 && pos < J.Lin) ? apply_tin(xv[q], J.tin, mean, rstd) : 0.f;
    }
  }
  __syncwarp();
  float w[K];
#pragma unroll
  for (int t = 0; t < K; ++t) w[t] = J.wd[c * K + t];
  float* ur = J.u + (long long)b * J.ubs + (long long)c * J.Lout;
  const bool vecok = (J.Lout % 4 == 0) && (J.ubs % 4 == 0);
  float* orow = outs[warp];
  for (int g = lane; l0 + g * 4 < l1; g += 32) {
    const int l = l0 + g * 4;
    float xw[NX4 * 4];
    const float4* xp = reinterpret_cast<const float4*>(row + g * 4 * S);
#pragma unroll
    for (int q = 0; q < NX4; ++q) { float4 v = xp[q]; xw[4 * q] = v.x; xw[4 * q + 1] = v.y; xw[4 * q + 2] = v.z; xw[4 * q + 3] = v.w; }
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < K; ++t)
#pragma unroll
      for (int n = 0; n < 4; ++n) a[n] = fmaf(w[t], xw[n * S + t * D], a[n]);
    if (vecok) *reinterpret_cast<float4*>(ur + l) = make_float4(a[0], a[1], a[2], a[3]);     // Lout % 4 == 0: l + 3 < Lout
    else *reinterpret_cast<float4*>(orow + g * 4) = make_float4(a[0], a[1], a[2], a[3]);
  }
  if (!vecok) {
    // rows that are not 16-byte aligned (L = 250, 125, 42): lane-interleaved stores, one full line per warp request
    __syncwarp();
    _Pragma("unroll 4")
    for (int l = l0 + lane; l < l1; l += 32) ur[l] = orow[l - l0];
  }
}

// second depthwise convolution of sep_conv_15 and sep_conv_9 in one launch (per-job kernel size)
static __global__ void __launch_bounds__(kThreads) k_dw2_fwd(const __grid_constant__ DwBatch P) {
  if (P.j[blockIdx.z].K == 15) dw_fwd_body<15, 1, 1>(P); else dw_fwd_body<9, 1, 1>(P);
}

// ------------------------------------------------------------------ fused edge forward (north star (1))
// The first depthwise convolution of every enabled sep_conv / dil_conv candidate of one edge in ONE pass:
// relu(x) is staged once per (row, segment) tile and each candidate writes its own output (was: one launch per
// candidate, each re-reading and re-transforming x).
struct EdgeFwdJob {
  const float* x; long long xbs;      // edge input [B][C][Lin]
  const float* wd[4];                 // [C][K], order {sep15, sep9, dil15, dil9}; null = candidate off
  float* u[4]; long long ubs;         // outputs [B][C][Lout]
  int Lin, Lout;
};
constexpr int kMaxEdgeFwd = 8;
struct EdgeFwdBatch { int n, B, C, nseg; EdgeFwdJob j[kMaxEdgeFwd]; };
template <int S>
struct EdgeFwdGeom {
  static constexpr int SEGO = S == 1 ? 256 : (S == 2 ? 128 : 84);
  static constexpr int XO = 16;                                                  // row: index p <-> pos = l0*S - XO + p
  static constexpr int WP = (SEGO * S + 2 * XO + 3 * S + 8 + 3) / 4 * 4;
};

template <int K, int D, int S>
__device__ __forceinline__ void edge_fwd_cand(const float* row, const float* wdp, float* ur, float* orow, int l0, int l1,
                                              int Lout, bool vecok) {
  using G = EdgeFwdGeom<S>;
  constexpr int PADC = (K - 1) * D / 2;
  constexpr int XS = G::XO - PADC, XA = XS / 4 * 4, XR = XS - XA;
  constexpr int NX4 = (3 * S + (K - 1) * D + 1 + XR + 3) / 4;
  const int lane = threadIdx.x & 31;
  float w[K];
#pragma unroll
  for (int t = 0; t < K; ++t) w[t] = wdp[t];
  __syncwarp();                                   // the previous candidate's staged outputs have been stored
  for (int g = lane; l0 + g * 4 < l1; g += 32) {
    const int l = l0 + g * 4;
    float xw[NX4 * 4];
    const float4* xp = reinterpret_cast<const float4*>(row + g * 4 * S + XA);
#pragma unroll
    for (int q = 0; q < NX4; ++q) { float4 v = xp[q]; xw[4 * q] = v.x; xw[4 * q + 1] = v.y; xw[4 * q + 2] = v.z; xw[4 * q + 3] = v.w; }
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < K; ++t)
#pragma unroll
      for (int n = 0; n < 4; ++n) a[n] = fmaf(w[t], xw[XR + n * S + t * D], a[n]);
    if (vecok) *reinterpret_cast<float4*>(ur + l) = make_float4(a[0], a[1], a[2], a[3]);
    else *reinterpret_cast<float4*>(orow + g * 4) = make_float4(a[0], a[1], a[2], a[3]);
  }
  if (!vecok) {
    __syncwarp();
    _Pragma("unroll 4")
    for (int l = l0 + lane; l < l1; l += 32) ur[l] = orow[l - l0];
  }
  (void)Lout;
}

// CTA: batch row b, 8 channels (one per warp), one output segment
template <int S>
__global__ void __launch_bounds__(kThreads) k_edge_fwd(const __grid_constant__ EdgeFwdBatch P) {
  using G = EdgeFwdGeom<S>;
  constexpr int SEGO = G::SEGO, WP = G::WP, XO = G::XO;
  const EdgeFwdJob& J = P.j[blockIdx.z];
  const int b = blockIdx.y / P.nseg, seg = blockIdx.y - b * P.nseg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 8 + warp;
  const int l0 = seg * SEGO, l1 = min(J.Lout, l0 + SEGO);
  __shared__ __align__(16) float rows[8][WP];
  __shared__ __align__(16) float outs[8][SEGO];
  if (c >= P.C || l0 >= J.Lout) return;
  float* row = rows[warp];
  {
    const float* xr = J.x + (long long)b * J.xbs + (long long)c * J.Lin;
    float xv[(WP + 31) / 32];
#pragma unroll
    for (int q = 0; q < (WP + 31) / 32; ++q) {
      const int pos = l0 * S - XO + lane + 32 * q;
      xv[q] = (pos >= 0 && pos < J.Lin) ? xr[pos] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < (WP + 31) / 32; ++q) {
      const int p = lane + 32 * q;
      if (p < WP) row[p] = xv[q] > 0.f ? xv[q] : 0.f;
    }
  }
  __syncwarp();
  const bool vecok = (J.Lout % 4 == 0) && (J.ubs % 4 == 0);
  const long long uo = (long long)b * J.ubs + (long long)c * J.Lout;
  if (J.wd[0]) edge_fwd_cand<15, 1, S>(row, J.wd[0] + c * 15, J.u[0] + uo, outs[warp], l0, l1, J.Lout, vecok);
  if (J.wd[1]) edge_fwd_cand<9, 1, S>(row, J.wd[1] + c * 9, J.u[1] + uo, outs[warp], l0, l1, J.Lout, vecok);
  if (J.wd[2]) edge_fwd_cand<15, 2, S>(row, J.wd[2] + c * 15, J.u[2] + uo, outs[warp], l0, l1, J.Lout, vecok);
  if (J.wd[3]) edge_fwd_cand<9, 2, S>(row, J.wd[3] + c * 9, J.u[3] + uo, outs[warp], l0, l1, J.Lout, vecok);
}

// ------------------------------------------------------------------ pools
struct PoolJob {
  const float* x; long long xbs;
  float* zmax; float* zavg;          // [B][C][Lout] contiguous, either may be null
  double* acc_max; double* acc_avg;  // sum | sumsq
  int Lin, Lout;
};
constexpr int kMaxPool = 8;
struct PoolBatch { int n, B, C, S, nseg; PoolJob j[kMaxPool]; };
constexpr int kPoolFwdSeg = 256;      // outputs per CTA row segment

// MaxPool1d(5,pad 2, -inf padding) and AvgPool1d(5,pad 2,count_include_pad=False) in one pass over x.
// CTA: channel c (blockIdx.x), 8 batch rows (one per warp).
static __global__ void __launch_bounds__(kThreads) k_pool_fwd(const __grid_constant__ PoolBatch P) {
  const PoolJob& J = P.j[blockIdx.z];
  const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rg = blockIdx.y / P.nseg, seg = blockIdx.y - rg * P.nseg;
  const int b = rg * 8 + warp;
  const int l0 = seg * kPoolFwdSeg, l1 = min(J.Lout, l0 + kPoolFwdSeg);
  __shared__ float red[4][8];
  float sm = 0.f, qm = 0.f, sa = 0.f, qa = 0.f;
  if (b < P.B) {
    const float* xr = J.x + (long long)b * J.xbs + (long long)c * J.Lin;
    const long long o = ((long long)b * P.C + c) * J.Lout;
    _Pragma("unroll 8")
    for (int l = l0 + lane; l < l1; l += 32) {
      const int s0 = l * P.S - 2;
      float mx = -INFINITY, sum = 0.f;
      int cnt = 0;
#pragma unroll
      for (int t = 0; t < 5; ++t) {
        const int pos = s0 + t;
        if (pos >= 0 && pos < J.Lin) { const float v = xr[pos]; mx = v > mx ? v : mx; sum += v; ++cnt; }
      }
      if (J.zmax) { J.zmax[o + l] = mx; sm += mx; qm += mx * mx; }
      if (J.zavg) { const float a = sum / (float)cnt; J.zavg[o + l] = a; sa += a; qa += a * a; }
    }
  }
  sm = warp_sum(sm); qm = warp_sum(qm); sa = warp_sum(sa); qa = warp_sum(qa);
  if (lane == 0) { red[0][warp] = sm; red[1][warp] = qm; red[2][warp] = sa; red[3][warp] = qa; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += (double)red[threadIdx.x][w];
    double* dst = threadIdx.x < 2 ? J.acc_max : J.acc_avg;
    if (dst) atomicAdd(&dst[(threadIdx.x & 1) * P.C + c], t);
  }
}

// ------------------------------------------------------------------ BN finalize
constexpr int kMaxSlots = 288;
struct FinBatch {
  double* acc;          // [slot][2][C]
  float* stat;          // [slot][2][C]  (mean | rstd)
  float* running;       // flat running-stat buffer (mean at off, var at off + C) or null
  int C, nslots, training;
  float momentum;
  short slot[kMaxSlots];
  float count[kMaxSlots];
  long long rs_off[kMaxSlots];
};

static __global__ void k_bn_finalize(const __grid_constant__ FinBatch P) {
  const int s = P.slot[blockIdx.x];
  const double n = (double)P.count[blockIdx.x];
  for (int c = threadIdx.x; c < P.C; c += blockDim.x) {
    const long long ro = P.rs_off[blockIdx.x];
    float* st = P.stat + (long long)s * 2 * P.C;
    if (P.training) {
      const double* a = P.acc + (long long)s * 2 * P.C;
      const double mean = a[c] / n;
      double var = a[P.C + c] / n - mean * mean;
      if (var < 0.0) var = 0.0;
      st[c] = (float)mean;
      st[P.C + c] = (float)(1.0 / sqrt(var + (double)kBnEps));
      if (P.running != nullptr && ro >= 0) {
        float* rm = P.running + ro;
        const double unb = n > 1.0 ? var * n / (n - 1.0) : var;
        rm[c] = (1.f - P.momentum) * rm[c] + P.momentum * (float)mean;
        rm[P.C + c] = (1.f - P.momentum) * rm[P.C + c] + P.momentum * (float)unb;
      }
    } else {
      const float* rm = P.running + ro;
      st[c] = rm[c];
      st[P.C + c] = 1.0f / sqrtf(rm[P.C + c] + kBnEps);
    }
  }
}

// ------------------------------------------------------------------ y = BN(z) [* gamma + beta]
struct BnApplyJob { const float* z; float* y; const float* stat; const float* gamma; const float* beta; int C, L; long long total; };
struct BnApplyBatch { int n; BnApplyJob j[4]; };

static __global__ void __launch_bounds__(kThreads) k_bn_apply(const __grid_constant__ BnApplyBatch P) {
  const BnApplyJob& J = P.j[blockIdx.y];
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < J.total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i / J.L) % J.C);
    float v = (J.z[i] - J.stat[c]) * J.stat[J.C + c];
    if (J.gamma) v = v * J.gamma[c] + J.beta[c];
    J.y[i] = v;
  }
}

// ------------------------------------------------------------------ alpha-weighted mix of one node
struct MixTerm {
  const float* z; long long bs;  // tensor [B][C][L] with batch stride
  const float* stat;             // mean | rstd, or null for the stride-1 skip (plain x)
  const float* mask;             // skip-connect dropout mask (pre-scaled) or null
  int widx;                      // index into the mixing-weight matrix
};
constexpr int kMaxTerms = 40;
struct MixBatch { int B, C, L, nterms; float* out; long long obs; const float* mixw; MixTerm t[kMaxTerms]; };

// node[b,c,l] = sum_terms w * (z - mean[c]) * rstd[c]      (model_discCNN.py:93,195)
// Thread = 4 consecutive positions of one row (128-bit loads when every row is 16-byte aligned); the term loop is
// unrolled so that four independent tensor streams are in flight per thread.
static __global__ void __launch_bounds__(kThreads) k_mix_fwd(const __grid_constant__ MixBatch P) {
  const int CL = P.C * P.L;
  const int L4 = (P.L + 3) / 4;                                  // groups of 4 positions per row
  const long long groups = (long long)P.B * P.C * L4;
  bool vec = (P.L % 4 == 0) && (P.obs % 4 == 0);
  for (int t = 0; t < P.nterms; ++t) vec = vec && (P.t[t].bs % 4 == 0);
  for (long long gi = (long long)blockIdx.x * blockDim.x + threadIdx.x; gi < groups; gi += (long long)gridDim.x * blockDim.x) {
    const long long rowi = gi / L4;
    const int l0 = (int)(gi - rowi * L4) * 4;
    const int b = (int)(rowi / P.C), c = (int)(rowi - (long long)b * P.C);
    const int r = c * P.L + l0;
    const int nv = min(4, P.L - l0);
    float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int t = 0; t < P.nterms; ++t) {
      const MixTerm& T = P.t[t];
      const float* zp = T.z + (long long)b * T.bs + r;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec) { const float4 q = *reinterpret_cast<const float4*>(zp); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
      else { for (int e = 0; e < nv; ++e) v[e] = zp[e]; }
      const float w = P.mixw[T.widx];
      if (T.stat) {
        const float rstd = T.stat[P.C + c], mean = T.stat[c];
#pragma unroll
        for (int e = 0; e < 4; ++e) a[e] = fmaf(w, (v[e] - mean) * rstd, a[e]);
      } else if (T.mask) {
        const float* mk = T.mask + (long long)b * CL + r;
        for (int e = 0; e < nv; ++e) a[e] = fmaf(w, v[e] * mk[e], a[e]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) a[e] = fmaf(w, v[e], a[e]);
      }
    }
    float* op = P.out + (long long)b * P.obs + r;
    if (vec) *reinterpret_cast<float4*>(op) = make_float4(a[0], a[1], a[2], a[3]);
    else for (int e = 0; e < nv; ++e) op[e] = a[e];
  }
}

}  // namespace gnas

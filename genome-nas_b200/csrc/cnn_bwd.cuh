// Backward kernels of the MixedOp / CNN cell path (fp32).  The identities used
// are those autograd derives for the reference modules:
// for a candidate ending in BN without affine, with upstream gradient g,
//   S1_c = sum g, S2_c = sum g*zhat,  dmix_k = sum_c S2_c,
//   dz   = (w/sigma_c) * (g - S1_c/n - zhat*S2_c/n)  =  A_c*g + Bz_c*z + D_c
// so every consumer reads (g, z) and three per-channel coefficients.
#pragma once
#include "common.cuh"
#include "cnn_fwd.cuh"

namespace gnas {

// ------------------------------------------------------------------ reductions for BN backward / dmix
struct RedTerm {
  const float* z; long long bs;   // forward tensor (raw pre-BN output, or x for the identity skip)
  const float* mask;              // dropout mask of the identity skip or null
  int slot;                       // backward accumulator slot, or -1 for the identity skip
  int widx;                       // mixing weight index (identity skip: dmix accumulates directly)
};
struct RedBatch {
  int B, C, L, nterms, s1_slot;   // s1_slot: where sum(g) goes (shared by all terms), -1: per term slot only
  const float* g; long long gbs;
  double* bacc;                   // [slot][2][C]
  double* dmix;                   // [n_edges*n_prims]
  RedTerm t[kMaxTerms];
};

// CTA: channel c, 8 batch rows (one per warp), kRedTerms terms (blockIdx.z).  bacc[slot] = (sum g, sum g*z).
// The terms of a chunk are reduced together: g is read once per position and the kRedTerms independent z streams
// keep that many loads in flight per lane; the term chunks give small-C cells enough CTAs to fill the machine.
constexpr int kRedTerms = 4;
static __global__ void __launch_bounds__(kThreads) k_bwd_reduce(const __grid_constant__ RedBatch P) {
  const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.y * 8 + warp;
  const int t0 = blockIdx.z * kRedTerms;
  const int nt = P.nterms - t0 < kRedTerms ? P.nterms - t0 : kRedTerms;     // may be <= 0 for the S1-only CTA
  __shared__ float red[8][kRedTerms + 1];
  const bool ok = b < P.B;
  const float* gr = P.g + (long long)(ok ? b : 0) * P.gbs + (long long)c * P.L;
  const float* zr[kRedTerms]; const float* mr[kRedTerms];
#pragma unroll
  for (int k = 0; k < kRedTerms; ++k) {
    const RedTerm& T = P.t[k < nt ? t0 + k : (t0 < P.nterms ? t0 : 0)];
    zr[k] = T.z + (long long)(ok ? b : 0) * T.bs + (long long)c * P.L;
    mr[k] = T.mask ? T.mask + ((long long)(ok ? b : 0) * P.C + c) * P.L : nullptr;
  }
  float s1 = 0.f, s2[kRedTerms];
#pragma unroll
  for (int k = 0; k < kRedTerms; ++k) s2[k] = 0.f;
  // rows that are 16-byte aligned (L % 4 == 0 and every batch stride a multiple of 4): 128-bit loads, 4 positions per lane
  bool vec = (P.L % 4 == 0) && (P.gbs % 4 == 0) && (reinterpret_cast<uintptr_t>(P.g) & 15) == 0;
#pragma unroll
  for (int k = 0; k < kRedTerms; ++k) {
    const RedTerm& T = P.t[k < nt ? t0 + k : (t0 < P.nterms ? t0 : 0)];
    vec = vec && (T.bs % 4 == 0) && T.mask == nullptr && (reinterpret_cast<uintptr_t>(T.z) & 15) == 0;
  }
  if (ok && P.nterms > 0 && vec) {
    _Pragma("unroll 2")
    for (int l = 4 * lane; l < P.L; l += 128) {
      const float4 g4 = *reinterpret_cast<const float4*>(gr + l);
      s1 += (g4.x + g4.y) + (g4.z + g4.w);
#pragma unroll
      for (int k = 0; k < kRedTerms; ++k) {
        const float4 z4 = *reinterpret_cast<const float4*>(zr[k] + l);
        s2[k] = fmaf(g4.x, z4.x, s2[k]); s2[k] = fmaf(g4.y, z4.y, s2[k]);
        s2[k] = fmaf(g4.z, z4.z, s2[k]); s2[k] = fmaf(g4.w, z4.w, s2[k]);
      }
    }
  } else if (ok && P.nterms > 0) {
    _Pragma("unroll 2")
    for (int l = lane; l < P.L; l += 32) {
      const float gl = gr[l];
      s1 += gl;
#pragma unroll
      for (int k = 0; k < kRedTerms; ++k) {
        float zv = zr[k][l];
        if (mr[k]) zv *= mr[k][l];
        s2[k] = fmaf(gl, zv, s2[k]);
      }
    }
  } else if (ok) {
    for (int l = lane; l < P.L; l += 32) s1 += gr[l];
  }
  s1 = warp_sum(s1);
#pragma unroll
  for (int k = 0; k < kRedTerms; ++k) s2[k] = warp_sum(s2[k]);
  if (lane == 0) {
    red[warp][kRedTerms] = s1;
#pragma unroll
    for (int k = 0; k < kRedTerms; ++k) red[warp][k] = s2[k];
  }
  __syncthreads();
  if (threadIdx.x <= kRedTerms) {
    const int k = threadIdx.x;
    double S = 0.0;
    for (int w = 0; w < 8; ++w) S += (double)red[w][k];
    if (k == kRedTerms) {
      if (blockIdx.z == 0 && P.s1_slot >= 0) atomicAdd(&P.bacc[(long long)P.s1_slot * 2 * P.C + c], S);
    } else if (k < nt) {
      const RedTerm& T = P.t[t0 + k];
      if (T.slot >= 0) atomicAdd(&P.bacc[(long long)T.slot * 2 * P.C + P.C + c], S);
      else atomicAdd(&P.dmix[T.widx], S);
    }
    // per-term sum(g) (no shared slot): the value sits in red[..][kRedTerms]
    if (k < nt && P.s1_slot < 0 && P.t[t0 + k].slot >= 0) {
      double S1 = 0.0;
      for (int w = 0; w < 8; ++w) S1 += (double)red[w][kRedTerms];
      atomicAdd(&P.bacc[(long long)P.t[t0 + k].slot * 2 * P.C + c], S1);
    }
  }
}
static inline int red_term_chunks(int nterms) { return nterms > 0 ? (nterms + kRedTerms - 1) / kRedTerms : 1; }

// ------------------------------------------------------------------ coefficients A | Bz | D
struct CoefBatch {
  const double* bacc; const float* stat; float* coef;   // indexed by slot
  const float* mixw; double* dmix;
  const float* gamma; float* dgamma; float* dbeta;      // affine BN (stem) or null
  int C, nslots, s1_slot;                               // s1_slot >= 0: sum(g) lives in that slot
  short slot[kMaxSlots]; short widx[kMaxSlots]; float count[kMaxSlots];
};

static __global__ void k_bn_bwd_coef(const __grid_constant__ CoefBatch P) {
  const int s = P.slot[blockIdx.x];
  const int wi = P.widx[blockIdx.x];
  const double n = (double)P.count[blockIdx.x];
  const double w = wi >= 0 ? (double)P.mixw[wi] : 1.0;
  const double* a = P.bacc + (long long)s * 2 * P.C;
  const double* a1 = P.s1_slot >= 0 ? P.bacc + (long long)P.s1_slot * 2 * P.C : a;
  const float* st = P.stat + (long long)s * 2 * P.C;
  float* cf = P.coef + (long long)s * 3 * P.C;
  __shared__ double red[32];
  double part = 0.0;
  for (int c = threadIdx.x; c < P.C; c += blockDim.x) {
    const double mean = st[c], rstd = st[P.C + c];
    const double S1 = a1[c];
    const double S2 = rstd * (a[P.C + c] - mean * S1);
    const double wc = P.gamma ? w * (double)P.gamma[c] : w;
    cf[c] = (float)(wc * rstd);
    cf[P.C + c] = (float)(-wc * rstd * rstd * S2 / n);
    cf[2 * P.C + c] = (float)(-wc * rstd * S1 / n + wc * rstd * rstd * mean * S2 / n);
    if (P.dgamma) { P.dgamma[c] += (float)S2; P.dbeta[c] += (float)S1; }
    part += S2;
  }
  part = warp_sum_d(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0 && wi >= 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x + 31) / 32; ++i) t += red[i];
    atomicAdd(&P.dmix[wi], t);
  }
}

// ------------------------------------------------------------------ transposed dense conv (data gradient)
struct DxJob {
  const float* g; long long gbs;     // upstream gradient [B][Cout][Lo]
  const float* z; long long zbs;     // raw forward output of the BN being differentiated, or null
  const float* coef;                 // A | Bz | D  [3][Cout]
  const float* wt;                   // packed [Cout][KW][Cin]
  float* out; long long obs;         // [B][Cin][Lin]
  const float* xin; long long xibs;  // forward input: x (O_RELU_ACC) or z1 (O_BNRELU_STORE)
  const float* xin_stat;             // mean | rstd of the inner BN
  double* bacc;                      // (sum v, sum v*z1)[Cin]
  int Cout, Lo, Lin, omode;
};
struct DxBatch { int n, B, Cin, nq, pad, KC; DxJob j[kMaxDense]; };

// out[b,ci,m] (op)= sum_{co,t : (m+pad-t)%S==0} W[co,ci,t] * dz[b,co,(m+pad-t)/S]
template <int KW, int S, int TM>
__global__ void __launch_bounds__(kThreads) k_dense_bwd_dx(const __grid_constant__ DxBatch P) {
  const DxJob& J = P.j[blockIdx.z];
  const int nq = P.nq, TLm = 4 * S * nq, m0 = blockIdx.x * TLm;
  if (m0 >= J.Lin) return;
  const int b = blockIdx.y, Cin = P.Cin, KC = P.KC;
  const int ncig = Cin / TM;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cig = tid / (S * nq), rem = tid - cig * (S * nq);
  const int rho = rem / nq, q = rem - rho * nq;
  const bool active = cig < ncig;
  constexpr int NT = (KW + S - 1) / S;           // max taps per residue class
  constexpr int NXW = 4 + NT;                    // window (incl. +1 for the residue shift)
  constexpr int NX4 = (NXW + 3) / 4;
  const int t0 = (rho + P.pad) % S;
  const int nt = t0 < KW ? (KW - t0 + S - 1) / S : 0;
  const int dc = (rho + P.pad - t0) / S - P.pad / S;   // 0 or 1
  const int lbase = m0 / S + P.pad / S - (NT - 1);
  const int Wu = 4 * nq + NT + 1;
  const int Wp = (Wu + 3) / 4 * 4 + 4;
  // Same two-stage pipeline as the forward kernel: weights of the next chunk of KC output channels by
  // cp.async, (g, z) rows of the next chunk prefetched into registers, dz = A*g + Bz*z + D formed on store.
  constexpr int NPRE = 6;
  GNAS_DYN_SMEM(float, smem);
  const int ds_sz = KC * Wp, ws_sz = KC * KW * Cin;
  float* dsb = smem;                              // [2][KC][Wp]  dz tiles
  float* wsb = dsb + 2 * ds_sz;                   // [2][KC][KW][Cin]
  double* ssum = reinterpret_cast<double*>(wsb + 2 * ws_sz);   // [2][Cin], double: order-independent (see k_dense_fwd)
  for (int i = tid; i < 2 * Cin; i += kThreads) ssum[i] = 0.0;
  float acc[TM][4];
#pragma unroll
  for (int m = 0; m < TM; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n) acc[m][n] = 0.f;
  const int nchunks = J.Cout / KC;
  const bool prefetch = KC <= kThreads / 32 && Wp <= 32 * NPRE;
  auto load_w = [&](int c, int buf) {
    const float* src = J.wt + (long long)c * ws_sz;
    float* dst = wsb + buf * ws_sz;
    for (int i = 4 * tid; i < ws_sz; i += 4 * kThreads) gnas_cp_async16(dst + i, src + i);
    gnas_cp_async_commit();
  };
  auto load_d_sync = [&](int c, int buf) {
    float* ds = dsb + buf * ds_sz;
    for (int co = warp; co < KC; co += kThreads / 32) {
      const int ch = c * KC + co;
      const float* gr = J.g + (long long)b * J.gbs + (long long)ch * J.Lo;
      const float* zr = J.z ? J.z + (long long)b * J.zbs + (long long)ch * J.Lo : nullptr;
      float A = 1.f, Bz = 0.f, D = 0.f;
      if (zr) { A = J.coef[ch]; Bz = J.coef[J.Cout + ch]; D = J.coef[2 * J.Cout + ch]; }
      _Pragma("unroll 4")
      for (int p = lane; p < Wp; p += 32) {
        const int l = lbase + p;
        float v = 0.f;
        if (p < Wu && l >= 0 && l < J.Lo) v = zr ? fmaf(A, gr[l], fmaf(Bz, zr[l], D)) : gr[l];
        ds[co * Wp + p] = v;
      }
    }
  };
  float gpre[NPRE], zpre[NPRE];
  auto fetch_d = [&](int c) {
    if (warp < KC) {
      const int ch = c * KC + warp;
      const float* gr = J.g + (long long)b * J.gbs + (long long)ch * J.Lo;
      const float* zr = J.z ? J.z + (long long)b * J.zbs + (long long)ch * J.Lo : nullptr;
#pragma unroll
      for (int q = 0; q < NPRE; ++q) {
        const int p = lane + 32 * q, l = lbase + p;
        const bool in = p < Wu && l >= 0 && l < J.Lo;
        gpre[q] = in ? gr[l] : 0.f;
        zpre[q] = (in && zr) ? zr[l] : 0.f;
      }
    }
  };
  auto store_d = [&](int c, int buf) {
    if (warp < KC) {
      const int ch = c * KC + warp;
      float* ds = dsb + buf * ds_sz + warp * Wp;
      float A = 1.f, Bz = 0.f, D = 0.f;
      if (J.z) { A = J.coef[ch]; Bz = J.coef[J.Cout + ch]; D = J.coef[2 * J.Cout + ch]; }
#pragma unroll
      for (int q = 0; q < NPRE; ++q) {
        const int p = lane + 32 * q, l = lbase + p;
        if (p < Wp) ds[p] = (p < Wu && l >= 0 && l < J.Lo) ? fmaf(A, gpre[q], fmaf(Bz, zpre[q], D)) : 0.f;
      }
    }
  };
  load_w(0, 0);
  load_d_sync(0, 0);
  gnas_cp_async_wait_all();
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    const int cur = c & 1, nxt = cur ^ 1;
    const bool more = c + 1 < nchunks;
    if (more) { load_w(c + 1, nxt); if (prefetch) fetch_d(c + 1); }
    if (active && nt > 0) {
      const float* ds = dsb + cur * ds_sz;
      const float* ws = wsb + cur * ws_sz;
      for (int co = 0; co < KC; ++co) {
        float xr[NX4 * 4];
        if (S == 1) {
          const float4* xp = reinterpret_cast<const float4*>(ds + co * Wp + 4 * q);
#pragma unroll
          for (int k = 0; k < NX4; ++k) { float4 v = xp[k]; xr[4 * k] = v.x; xr[4 * k + 1] = v.y; xr[4 * k + 2] = v.z; xr[4 * k + 3] = v.w; }
        } else {
          const float* xp = ds + co * Wp + 4 * q + dc;
#pragma unroll
          for (int k = 0; k < NXW - 1; ++k) xr[k] = xp[k];
        }
        const float* wrow = ws + co * KW * Cin + cig * TM;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          if (t < nt) {
            float w[TM];
            WVec<TM>::ld(wrow + (t0 + S * t) * Cin, w);
#pragma unroll
            for (int m = 0; m < TM; ++m)
#pragma unroll
              for (int n = 0; n < 4; ++n) acc[m][n] = fmaf(w[m], xr[n + NT - 1 - t], acc[m][n]);
          }
        }
      }
    }
    if (more) {
      if (prefetch) store_d(c + 1, nxt); else load_d_sync(c + 1, nxt);
      gnas_cp_async_wait_all();
    }
    __syncthreads();
  }
  // ---- epilogue
  const bool pow2 = (nq & (nq - 1)) == 0;
  const int width = nq < 32 ? nq : 32;
#pragma unroll
  for (int m = 0; m < TM; ++m) {
    const int ci = cig * TM + m;
    float s = 0.f, qq = 0.f;
    if (active && (nt > 0 || J.omode == O_STORE || J.omode == O_BNRELU_STORE)) {
      float* orow = J.out + (long long)b * J.obs + (long long)ci * J.Lin;
      const float* xrow = J.xin ? J.xin + (long long)b * J.xibs + (long long)ci * J.Lin : nullptr;
      const float mean = (J.omode == O_BNRELU_STORE) ? J.xin_stat[ci] : 0.f;
      // all loads first (they are independent; interleaving them with the stores would serialise on aliasing)
      float xv[4], ov[4];
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int pos = m0 + S * (4 * q + n) + rho;
        const bool in = pos < J.Lin;
        xv[n] = (in && xrow) ? xrow[pos] : 0.f;
        ov[n] = (in && (J.omode == O_ACC || J.omode == O_RELU_ACC)) ? orow[pos] : 0.f;
      }
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int pos = m0 + S * (4 * q + n) + rho;
        if (pos < J.Lin) {
          const float v = acc[m][n];
          if (J.omode == O_STORE) orow[pos] = v;
          else if (J.omode == O_ACC) orow[pos] = ov[n] + v;
          else if (J.omode == O_RELU_ACC) { if (xv[n] > 0.f) orow[pos] = ov[n] + v; }
          else {
            const float vv = xv[n] > mean ? v : 0.f;
            orow[pos] = vv; s += vv; qq = fmaf(vv, xv[n], qq);
          }
        }
      }
    }
    if (J.omode == O_BNRELU_STORE) {
      if (pow2 && S == 1) {
        for (int o = width >> 1; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); qq += __shfl_xor_sync(0xffffffffu, qq, o); }
        if (active && (q & (width - 1)) == 0) { atomicAdd(&ssum[ci], (double)s); atomicAdd(&ssum[Cin + ci], (double)qq); }
      } else if (active) {
        atomicAdd(&ssum[ci], (double)s); atomicAdd(&ssum[Cin + ci], (double)qq);
      }
    }
  }
  if (J.omode == O_BNRELU_STORE) {
    __syncthreads();
    for (int i = tid; i < 2 * Cin; i += kThreads) atomicAdd(&J.bacc[i], ssum[i]);
  }
}

// ------------------------------------------------------------------ pointwise (1x1) weight gradient
struct WgJob {
  const float* g; long long gbs;     // upstream gradient [B][Cout][Lo]
  const float* z; long long zbs;     // raw BN input (affine dz) or null
  const float* coef;
  const float* r; long long rbs;     // forward input of the conv [B][Cin][Lin]
  const float* r_stat;               // for T_BN_RELU
  float* dw;                         // [Cout][Cin][KW], accumulated with atomics
  int Cin, Lo, Lin, tin;
};
struct WgBatch { int n, B, Cout, S, pad, TPC, TL, COB, CIB, G; WgJob j[kMaxDense]; };

// Position tiles are enumerated b-major: tile -> (b, l0).  A CTA owns TPC consecutive tiles (blockIdx.y),
// one block of the weight matrix (blockIdx.x) of one job (blockIdx.z), accumulates in registers and
// flushes once with atomics into the flat gradient buffer.

// dW[co,ci] += sum_{b,l} dz[b,co,l] * T(r)[b,ci,l*S]
template <int TC>
__global__ void __launch_bounds__(kThreads) k_pw_wgrad(const __grid_constant__ WgBatch P) {
  const WgJob& J = P.j[blockIdx.z];
  const int nbi = (J.Cin + P.CIB - 1) / P.CIB;
  const int wb = blockIdx.x;
  const int co0 = (wb / nbi) * P.COB, ci0 = (wb % nbi) * P.CIB;
  const int cob = P.COB, cib = P.CIB;
  const int nci_t = cib / TC;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int per = (cob / TC) * nci_t;
  const int G = kThreads / per;
  const int grp = tid / per, r2 = tid - grp * per;
  const int cq = r2 / nci_t, iq = r2 - cq * nci_t;
  const bool active = grp < G;
  const int TL = P.TL;
  const int po = cob + 4, pi = cib + 4;
  GNAS_DYN_SMEM(float, smem);
  float* dzs = smem;             // [TL][po]
  float* rs = dzs + TL * po;     // [TL][pi]
  float acc[TC][TC];
#pragma unroll
  for (int i = 0; i < TC; ++i)
#pragma unroll
    for (int k = 0; k < TC; ++k) acc[i][k] = 0.f;
  const int ntl = (J.Lo + TL - 1) / TL;
  const int tile_end = min(P.B * ntl, (int)(blockIdx.y + 1) * P.TPC);
  for (int tile = blockIdx.y * P.TPC; tile < tile_end; ++tile) {
    const int b = tile / ntl, l0 = (tile - b * ntl) * TL;
    const int nl = min(TL, J.Lo - l0);
    for (int c = warp; c < cob; c += kThreads / 32) {
      const long long o = (long long)b * J.gbs + (long long)(co0 + c) * J.Lo + l0;
      const float* gr = J.g + o;
      if (J.z) {
        const float* zr = J.z + (long long)b * J.zbs + (long long)(co0 + c) * J.Lo + l0;
        const float A = J.coef[co0 + c], Bz = J.coef[P.Cout + co0 + c], D = J.coef[2 * P.Cout + co0 + c];
        _Pragma("unroll 4")
        for (int l = lane; l < TL; l += 32) dzs[l * po + c] = l < nl ? fmaf(A, gr[l], fmaf(Bz, zr[l], D)) : 0.f;
      } else {
        _Pragma("unroll 4")
        for (int l = lane; l < TL; l += 32) dzs[l * po + c] = l < nl ? gr[l] : 0.f;
      }
    }
    for (int c = warp; c < cib; c += kThreads / 32) {
      const float* rr = J.r + (long long)b * J.rbs + (long long)(ci0 + c) * J.Lin + (long long)l0 * P.S;
      float mean = 0.f, rstd = 1.f;
      if (J.tin >= T_BN_RELU) { mean = J.r_stat[ci0 + c]; rstd = J.r_stat[J.Cin + ci0 + c]; }
      _Pragma("unroll 4")
      for (int l = lane; l < TL; l += 32) rs[l * pi + c] = l < nl ? apply_tin(rr[l * P.S], J.tin, mean, rstd) : 0.f;
    }
    __syncthreads();
    if (active) {
      const float* dp = dzs + cq * TC;
      const float* rp = rs + iq * TC;
      for (int l = grp; l < nl; l += G) {
        float a[TC], bb[TC];
        if (TC >= 4) {
#pragma unroll
          for (int i = 0; i < TC; i += 4) {
            const float4 v = *reinterpret_cast<const float4*>(dp + l * po + i);
            a[i] = v.x; a[i + 1 < TC ? i + 1 : 0] = v.y; a[i + 2 < TC ? i + 2 : 0] = v.z; a[i + 3 < TC ? i + 3 : 0] = v.w;
            const float4 u = *reinterpret_cast<const float4*>(rp + l * pi + i);
            bb[i] = u.x; bb[i + 1 < TC ? i + 1 : 0] = u.y; bb[i + 2 < TC ? i + 2 : 0] = u.z; bb[i + 3 < TC ? i + 3 : 0] = u.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < TC; ++i) { a[i] = dp[l * po + i]; bb[i] = rp[l * pi + i]; }
        }
#pragma unroll
        for (int i = 0; i < TC; ++i)
#pragma unroll
          for (int k = 0; k < TC; ++k) acc[i][k] = fmaf(a[i], bb[k], acc[i][k]);
      }
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < TC; ++i)
#pragma unroll
      for (int k = 0; k < TC; ++k)
        atomicAdd(&J.dw[(long long)(co0 + cq * TC + i) * J.Cin + ci0 + iq * TC + k], acc[i][k]);
  }
}

// dW[co,ci,t] += sum_{b,l} dz[b,co,l] * T(r)[b,ci,l*S+t-pad]     (KW = 15 cells, 13 stem)
template <int KW, int S>
__global__ void __launch_bounds__(kThreads) k_conv_wgrad(const __grid_constant__ WgBatch P) {
  const WgJob& J = P.j[blockIdx.z];
  const int nbi = (J.Cin + P.CIB - 1) / P.CIB;
  const int wb = blockIdx.x;
  const int co0 = (wb / nbi) * P.COB, ci0 = (wb % nbi) * P.CIB;
  const int cob = P.COB, cib = P.CIB;
  const int ncoq = cob / 4;
  const int per = ncoq * cib;
  const int G = kThreads / per;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int grp = tid / per, r2 = tid - grp * per;
  const int coq = r2 / cib, ci = r2 - coq * cib;
  const bool active = grp < G;
  const int TLg = P.TL;                 // positions per group per tile (multiple of 4)
  const int TLc = TLg * G;              // positions per tile
  const int pz = (TLc + 3) / 4 * 4 + 4;  // multiple of 4: 128-bit (warp-broadcast) loads of dz
  const int Wr = (TLc - 1) * S + KW;
  const int pr = Wr | 1;                // odd: lanes (= input channels) hit distinct banks
  GNAS_DYN_SMEM(float, smem);
  float* dzs = smem;                    // [cob][pz]
  float* rs = dzs + cob * pz;           // [cib][pr]
  float acc[4][KW];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int t = 0; t < KW; ++t) acc[i][t] = 0.f;
  const int ntl = (J.Lo + TLc - 1) / TLc;
  const int tile_end = min(P.B * ntl, (int)(blockIdx.y + 1) * P.TPC);
  for (int tile = blockIdx.y * P.TPC; tile < tile_end; ++tile) {
    const int b = tile / ntl, l0 = (tile - b * ntl) * TLc;
    const int nl = min(TLc, J.Lo - l0);
    for (int c = warp; c < cob; c += kThreads / 32) {
      const float* gr = J.g + (long long)b * J.gbs + (long long)(co0 + c) * J.Lo + l0;
      if (J.z) {
        const float* zr = J.z + (long long)b * J.zbs + (long long)(co0 + c) * J.Lo + l0;
        const float A = J.coef[co0 + c], Bz = J.coef[P.Cout + co0 + c], D = J.coef[2 * P.Cout + co0 + c];
        _Pragma("unroll 4")
        for (int l = lane; l < TLc; l += 32) dzs[c * pz + l] = l < nl ? fmaf(A, gr[l], fmaf(Bz, zr[l], D)) : 0.f;
      } else {
        _Pragma("unroll 4")
        for (int l = lane; l < TLc; l += 32) dzs[c * pz + l] = l < nl ? gr[l] : 0.f;
      }
    }
    const int in0 = l0 * S - P.pad;
    for (int c = warp; c < cib; c += kThreads / 32) {
      const float* rr = J.r + (long long)b * J.rbs + (long long)(ci0 + c) * J.Lin;
      float mean = 0.f, rstd = 1.f;
      if (J.tin >= T_BN_RELU) { mean = J.r_stat[ci0 + c]; rstd = J.r_stat[J.Cin + ci0 + c]; }
      _Pragma("unroll 4")
      for (int p = lane; p < Wr; p += 32) {
        const int pos = in0 + p;
        rs[c * pr + p] = (pos >= 0 && pos < J.Lin) ? apply_tin(rr[pos], J.tin, mean, rstd) : 0.f;
      }
    }
    __syncthreads();
    if (active) {
      // 4 positions per step: 4 x 128-bit dz loads (broadcast) + one (3S+KW)-wide window of r -> 16*KW FMAs
      constexpr int NXW = 3 * S + KW;
      const int le = min(nl, (grp + 1) * TLg);
      const float* rr = rs + ci * pr;
      const float* dz0 = dzs + (coq * 4) * pz;
      for (int l = grp * TLg; l < le; l += 4) {
        float d[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(dz0 + i * pz + l);
          d[i][0] = v.x; d[i][1] = v.y; d[i][2] = v.z; d[i][3] = v.w;
        }
        float rw[NXW];
        const float* rl = rr + l * S;
#pragma unroll
        for (int k = 0; k < NXW; ++k) rw[k] = rl[k];
#pragma unroll
        for (int t = 0; t < KW; ++t)
#pragma unroll
          for (int n = 0; n < 4; ++n) {
            const float rv = rw[n * S + t];
            acc[0][t] = fmaf(d[0][n], rv, acc[0][t]); acc[1][t] = fmaf(d[1][n], rv, acc[1][t]);
            acc[2][t] = fmaf(d[2][n], rv, acc[2][t]); acc[3][t] = fmaf(d[3][n], rv, acc[3][t]);
          }
      }
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int t = 0; t < KW; ++t)
        atomicAdd(&J.dw[((long long)(co0 + coq * 4 + i) * J.Cin + ci0 + ci) * KW + t], acc[i][t]);
  }
}

// ------------------------------------------------------------------ depthwise backward
struct DwBwdJob {
  const float* du; long long dubs;   // gradient wrt the depthwise output [B][C][Lo]
  const float* x; long long xbs;     // forward input [B][C][Lin]
  const float* in_stat;              // inner BN mean|rstd (T_BN_RELU)
  const float* wd;                   // [C][K]
  float* out; long long obs;         // dX accumulate (O_RELU_ACC) or dyh store (O_BNRELU_STORE)
  double* bacc;                      // (sum v, sum v*z1)[C] for O_BNRELU_STORE
  float* dwd;                        // [C][K] atomics, or null
  int Lin, Lo, tin, omode;
  int K;                             // 15 or 9 (k_dw2_bwd picks the instance per job)
};
struct DwBwdBatch { int n, B, C, pad, RPW, nseg; DwBwdJob j[kMaxDw]; };

// Tile geometry shared by the kernel and its launcher.  A row [0, Lin) is cut into segments of SEG input positions
// (a multiple of 4 and of every stride): with whole rows a launch of the C = 8 cells had 1.5 CTAs per SM and every
// warp walked three dependent phases over 1000 positions -- latency-bound at 9 % of the HBM roofline.
template <int K, int D, int S>
struct DwBwdGeom {
  static constexpr int SEG = S == 1 ? 256 : 252;      // 252 = 4 * 63: a multiple of 2, 3 and 4
  static constexpr int PADC = (K - 1) * D / 2;
  // du row: index i <-> l = l0 + i - OFF, OFF a multiple of 4 >= PADC (stride 1) so that windows start 16-byte aligned
  static constexpr int OFF = S == 1 ? (PADC + 3) / 4 * 4 : K;
  static constexpr int WX = (SEG + 2 * PADC + 8 + 3) / 4 * 4;                   // x row: index p <-> pos = m0 - pad + p
  static constexpr int WD = S == 1 ? (SEG + 2 * PADC + 16 + 3) / 4 * 4 : ((SEG + PADC) / S + OFF + 4 + 3) / 4 * 4;
};

// CTA: channel c, 8 batch rows (one per warp), one segment of SEG input positions.  Stride 1 (all but the first conv
// of strided edges): every lane produces 4 consecutive positions from register windows (128-bit shared loads, static
// tap indices).  The staging loops have compile-time bounds: all global loads of a tile are in flight at once.
template <int K, int D, int S>
__device__ __forceinline__ void dw_bwd_body(const DwBwdBatch& P) {
  using G = DwBwdGeom<K, D, S>;
  constexpr int PADC = G::PADC, OFF = G::OFF, SEG = G::SEG, WX = G::WX, WD = G::WD;
  static_assert(SEG % S == 0 && SEG % 4 == 0, "segment must be a multiple of the stride and of 4");
  const DwBwdJob& J = P.j[blockIdx.z];
  const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rg = blockIdx.y / P.nseg, seg = blockIdx.y - rg * P.nseg;
  const int m0 = seg * SEG;                               // first input position of the segment
  const int m1 = min(J.Lin, m0 + SEG);
  const int l0 = m0 / S;                                  // du positions l with l*S in [m0, m1)
  const int l1 = m1 > m0 ? min(J.Lo, (m1 - 1) / S + 1) : l0;
  GNAS_DYN_SMEM(float, smem);
  float* xrow = smem + warp * (WX + WD + SEG);   // T(x) with zero halo
  float* drow = xrow + WX;                     // du with zero halo
  float* arow = drow + WD;                     // dx of the segment before masking (stride 1)
  float mean = 0.f, rstd = 1.f;
  if (J.tin >= T_BN_RELU) { mean = J.in_stat[c]; rstd = J.in_stat[P.C + c]; }
  float w[K];
#pragma unroll
  for (int t = 0; t < K; ++t) w[t] = J.wd[c * K + t];
  float s = 0.f, qq = 0.f;
  float gw[K];
#pragma unroll
  for (int t = 0; t < K; ++t) gw[t] = 0.f;
  // each warp walks RPW batch rows of channel c and reduces its weight-gradient / BN sums once at the end
  for (int r = 0; r < P.RPW; ++r) {
  const int b = (rg * P.RPW + r) * 8 + warp;
  const bool ok = b < P.B && m0 < J.Lin;
  const float* xg = J.x + (long long)(ok ? b : 0) * J.xbs + (long long)c * J.Lin;
  __syncwarp();
  if (ok) {
    const float* dg = J.du + (long long)b * J.dubs + (long long)c * J.Lo;
    float xv_[(WX + 31) / 32], dv_[(WD + 31) / 32];
#pragma unroll
    for (int q = 0; q < (WX + 31) / 32; ++q) {
      const int pos = m0 - PADC + lane + 32 * q;
      xv_[q] = (pos >= 0 && pos < J.Lin) ? xg[pos] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < (WD + 31) / 32; ++q) {
      const int l = l0 - OFF + lane + 32 * q;
      dv_[q] = (l >= 0 && l < J.Lo) ? dg[l] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < (WX + 31) / 32; ++q) {
      const int p = lane + 32 * q, pos = m0 - PADC + p;
      if (p < WX) xrow[p] = (pos >= 0 && pos < J.Lin) ? apply_tin(xv_[q], J.tin, mean, rstd) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < (WD + 31) / 32; ++q) {
      const int i = lane + 32 * q;
      if (i < WD) drow[i] = dv_[q];
    }
  }
  __syncwarp();
  if (ok) {
    float* orow = J.out + (long long)b * J.obs + (long long)c * J.Lin;
    const int nm = m1 - m0, nl = l1 - l0;
    if (S == 1) {
      // dr[m] = sum_t w[t] * du[m + pad - t*D];  window base (drow index) = (m - m0) + pad - (K-1)*D + OFF = ml + SH
      constexpr int SH = OFF - PADC;                      // 0..3
      constexpr int NW = (K - 1) * D + 4 + SH;
      constexpr int NW4 = (NW + 3) / 4;
      for (int g4 = lane; g4 * 4 < nm; g4 += 32) {
        const int ml = g4 * 4;
        float dw_[NW4 * 4];
        const float4* dp = reinterpret_cast<const float4*>(drow + ml);
#pragma unroll
        for (int q = 0; q < NW4; ++q) { const float4 v = dp[q]; dw_[4 * q] = v.x; dw_[4 * q + 1] = v.y; dw_[4 * q + 2] = v.z; dw_[4 * q + 3] = v.w; }
        float a[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < K; ++t)
#pragma unroll
          for (int n = 0; n < 4; ++n) a[n] = fmaf(w[t], dw_[SH + n + (K - 1 - t) * D], a[n]);
        *reinterpret_cast<float4*>(arow + ml) = make_float4(a[0], a[1], a[2], a[3]);
      }
      __syncwarp();
      // the global read-modify-write walks the segment lane-interleaved: every access is one full 128-byte line per
      // warp whatever the alignment of the row (4 consecutive positions per lane cost 4 partial-sector requests each)
      _Pragma("unroll 4")
      for (int ml = lane; ml < nm; ml += 32) {
        const int m = m0 + ml;
        if (J.omode == O_RELU_ACC) { if (xrow[ml + PADC] > 0.f) orow[m] += arow[ml]; }
        else { const float xv = xg[m]; const float v = xv > mean ? arow[ml] : 0.f; orow[m] = v; s += v; qq = fmaf(v, xv, qq); }
      }
      // dwd[c,t] += sum_l du[l] * T(x)[l + t*D - pad]   (xrow index = (l - l0) + t*D)
      if (J.dwd) {
        constexpr int NR = (K - 1) * D + 4;
        constexpr int NR4 = (NR + 3) / 4;
        for (int g4 = lane; g4 * 4 < nl; g4 += 32) {
          const int ll = g4 * 4;
          const float4 dv = *reinterpret_cast<const float4*>(drow + ll + OFF);
          const float d4[4] = {dv.x, dv.y, dv.z, dv.w};
          float xw[NR4 * 4];
          const float4* xp = reinterpret_cast<const float4*>(xrow + ll);
#pragma unroll
          for (int q = 0; q < NR4; ++q) { const float4 v = xp[q]; xw[4 * q] = v.x; xw[4 * q + 1] = v.y; xw[4 * q + 2] = v.z; xw[4 * q + 3] = v.w; }
#pragma unroll
          for (int t = 0; t < K; ++t)
#pragma unroll
            for (int n = 0; n < 4; ++n) gw[t] = fmaf(d4[n], xw[n + t * D], gw[t]);
        }
      }
    } else {
      _Pragma("unroll 4")
      for (int ml = lane; ml < nm; ml += 32) {
        const int m = m0 + ml;
        float a = 0.f;
#pragma unroll
        for (int t = 0; t < K; ++t) {
          const int num = m + PADC - t * D;
          if (num >= 0 && num % S == 0) a = fmaf(w[t], drow[num / S - l0 + OFF], a);
        }
        const float xv = xg[m];
        if (J.omode == O_RELU_ACC) { if (xv > 0.f) orow[m] += a; }
        else { const float v = xv > mean ? a : 0.f; orow[m] = v; s += v; qq = fmaf(v, xv, qq); }
      }
      if (J.dwd) {
        _Pragma("unroll 4")
        for (int ll = lane; ll < nl; ll += 32) {
          const float dv = drow[ll + OFF];
#pragma unroll
          for (int t = 0; t < K; ++t) gw[t] = fmaf(dv, xrow[ll * S + t * D], gw[t]);
        }
      }
    }
  }
  }   // rows
  // CTA reduction of the K weight-gradient taps (weight pass only) and the two BN sums (stage-2 depthwise only)
  // through shared memory: 17 shuffle reductions per warp cost more instructions than the tile itself
  __syncthreads();                              // the row buffers are dead: reuse them as [K + 2][256]
  float* part = smem;
  const bool want_w = J.dwd != nullptr, want_s = J.omode == O_BNRELU_STORE;
  if (want_w) {
#pragma unroll
    for (int t = 0; t < K; ++t) part[t * kThreads + threadIdx.x] = gw[t];
  }
  if (want_s) { part[K * kThreads + threadIdx.x] = s; part[(K + 1) * kThreads + threadIdx.x] = qq; }
  __syncthreads();
  for (int v = warp; v < K + 2; v += 8) {
    if (v < K ? !want_w : !want_s) continue;
    const float* pv = part + v * kThreads;
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < kThreads / 32; ++q) t += pv[lane + 32 * q];
    t = warp_sum(t);
    if (lane == 0) {
      if (v < K) atomicAdd(&J.dwd[c * K + v], t);
      else atomicAdd(&J.bacc[(v - K) * P.C + c], (double)t);
    }
  }
}

// second depthwise convolution of sep_conv_15 and sep_conv_9 (stride 1, dilation 1) in one launch: the kernel size is
// a per-job switch between the two instances (two dependent launches per node before)
static __global__ void __launch_bounds__(kThreads, 4) k_dw2_bwd(const __grid_constant__ DwBwdBatch P) {
  if (P.j[blockIdx.z].K == 15) dw_bwd_body<15, 1, 1>(P); else dw_bwd_body<9, 1, 1>(P);
}

// ------------------------------------------------------------------ fused edge backward (north star (1): fused backward)
// All depthwise candidates that write the data gradient of one edge -- the first depthwise conv of sep_conv_15 /
// sep_conv_9 and the depthwise conv of dil_conv_15 / dil_conv_9 -- in ONE pass over the edge input: relu(x) is
// staged once, the four transposed depthwise convolutions accumulate in shared memory, and dX is read-modify-
// written once (was: four launches, each re-reading x and read-modify-writing dX: 16 tensor passes -> 7).
struct EdgeBwdJob {
  const float* du[4]; long long dubs;   // gradients wrt the depthwise outputs [B][C][Lo], order {sep15, sep9, dil15, dil9}; null = off
  const float* wd[4];                   // [C][K]
  float* dwd[4];                        // [C][K] atomics, or null
  const float* x; long long xbs;        // edge input [B][C][Lin] (raw; every candidate starts with ReLU)
  float* out; long long obs;            // dX, accumulated where x > 0
  int Lin, Lo;
};
constexpr int kMaxEdgeJobs = 8;
struct EdgeBwdBatch { int n, B, C, nseg; EdgeBwdJob j[kMaxEdgeJobs]; };

template <int S>
struct EdgeBwdGeom {
  static constexpr int SEG = S == 1 ? 256 : 252;
  static constexpr int XO = 16;                                    // x row: index p <-> pos = m0 - XO + p (XO >= largest pad, multiple of 4)
  static constexpr int WX = SEG + 2 * XO + 8;
  static constexpr int WD = S == 1 ? SEG + 2 * 14 + 16 : ((SEG + 14) / S + 15 + 4 + 3) / 4 * 4;
  static constexpr int ROW = WX + WD + SEG;                        // floats per warp
  static constexpr int PART = 15 * kThreads;                       // CTA reduction scratch of one candidate's taps
};

// one candidate of the fused edge backward: accumulate its dX into arow, its weight gradient into dwd
template <int K, int D, int S>
__device__ __forceinline__ void edge_bwd_cand(const float* dg, const float* wdp, float* dwdp, int Lo,
                                              const float* xrow, float* drow, float* arow, float* part,
                                              int m0, int m1, int l0, int l1, bool ok) {
  using G = EdgeBwdGeom<S>;
  constexpr int PADC = (K - 1) * D / 2;
  constexpr int OFF = S == 1 ? (PADC + 3) / 4 * 4 : K;            // du row: index i <-> l = l0 - OFF + i
  constexpr int WDV = S == 1 ? (G::SEG + 2 * PADC + 16 + 3) / 4 * 4 : ((G::SEG + PADC) / S + OFF + 4 + 3) / 4 * 4;
  static_assert(WDV <= G::WD, "du row does not fit");
  const int lane = threadIdx.x & 31;
  float w[K], gw[K];
#pragma unroll
  for (int t = 0; t < K; ++t) { w[t] = wdp[t]; gw[t] = 0.f; }
  __syncwarp();                                   // the previous candidate is done with drow
  if (ok) {
    float dv_[(WDV + 31) / 32];
#pragma unroll
    for (int q = 0; q < (WDV + 31) / 32; ++q) {
      const int l = l0 - OFF + lane + 32 * q;
      dv_[q] = (l >= 0 && l < Lo) ? dg[l] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < (WDV + 31) / 32; ++q) {
      const int i = lane + 32 * q;
      if (i < WDV) drow[i] = dv_[q];
    }
  }
  __syncwarp();
  if (ok) {
    const int nm = m1 - m0, nl = l1 - l0;
    constexpr int XS = G::XO - PADC;              // xrow index of position (l*S + t*D - pad) = (l - l0)*S + t*D + XS
    if (S == 1) {
      constexpr int SH = OFF - PADC;              // 0..3
      constexpr int NW4 = ((K - 1) * D + 4 + SH + 3) / 4;
      for (int g4 = lane; g4 * 4 < nm; g4 += 32) {
        const int ml = g4 * 4;
        float dw_[NW4 * 4];
        const float4* dp = reinterpret_cast<const float4*>(drow + ml);
#pragma unroll
        for (int q = 0; q < NW4; ++q) { const float4 v = dp[q]; dw_[4 * q] = v.x; dw_[4 * q + 1] = v.y; dw_[4 * q + 2] = v.z; dw_[4 * q + 3] = v.w; }
        float4 a = *reinterpret_cast<const float4*>(arow + ml);
#pragma unroll
        for (int t = 0; t < K; ++t) {
          a.x = fmaf(w[t], dw_[SH + 0 + (K - 1 - t) * D], a.x); a.y = fmaf(w[t], dw_[SH + 1 + (K - 1 - t) * D], a.y);
          a.z = fmaf(w[t], dw_[SH + 2 + (K - 1 - t) * D], a.z); a.w = fmaf(w[t], dw_[SH + 3 + (K - 1 - t) * D], a.w);
        }
        *reinterpret_cast<float4*>(arow + ml) = a;
      }
      if (dwdp) {
        constexpr int XA = XS / 4 * 4, XR = XS - XA;          // aligned window base + in-register shift
        constexpr int NR4 = ((K - 1) * D + 4 + XR + 3) / 4;
        for (int g4 = lane; g4 * 4 < nl; g4 += 32) {
          const int ll = g4 * 4;
          const float4 dv = *reinterpret_cast<const float4*>(drow + ll + OFF);
          const float d4[4] = {dv.x, dv.y, dv.z, dv.w};
          float xw[NR4 * 4];
          const float4* xp = reinterpret_cast<const float4*>(xrow + ll + XA);
#pragma unroll
          for (int q = 0; q < NR4; ++q) { const float4 v = xp[q]; xw[4 * q] = v.x; xw[4 * q + 1] = v.y; xw[4 * q + 2] = v.z; xw[4 * q + 3] = v.w; }
#pragma unroll
          for (int t = 0; t < K; ++t)
#pragma unroll
            for (int n = 0; n < 4; ++n) gw[t] = fmaf(d4[n], xw[XR + n + t * D], gw[t]);
        }
      }
    } else {
      // polyphase: position m receives the taps t with (m + pad - t D) % S == 0, i.e. t = t0, t0 + STEP, ... where
      // t0 < STEP = S / gcd(D, S) solves the congruence (none for D = S = 2 at odd m + pad): K / STEP multiply-adds per
      // position instead of K divisibility tests
      _Pragma("unroll 2")
      for (int ml = lane; ml < nm; ml += 32) {
        const int m = m0 + ml;
        constexpr int STEP = (D % S == 0) ? 1 : S;      // S / gcd(D, S) for D in {1, 2}, S in {2, 3}
        const int r = (m + PADC) % S;
        int t0 = -1;
#pragma unroll
        for (int c = STEP - 1; c >= 0; --c) if ((r + S * D - c * D) % S == 0) t0 = c;
        float a = arow[ml];
#pragma unroll
        for (int c = 0; c < STEP; ++c) {
          if (t0 != c) continue;            // static tap indices per residue: w[] stays in registers
#pragma unroll
          for (int u = 0; u < (K + STEP - 1) / STEP; ++u) {
            const int t = c + u * STEP;
            const int num = m + PADC - t * D;
            if (t < K && num >= 0) a = fmaf(w[t < K ? t : 0], drow[num / S - l0 + OFF], a);
          }
        }
        arow[ml] = a;
      }
      if (dwdp) {
        _Pragma("unroll 4")
        for (int ll = lane; ll < nl; ll += 32) {
          const float dv = drow[ll + OFF];
#pragma unroll
          for (int t = 0; t < K; ++t) gw[t] = fmaf(dv, xrow[ll * S + t * D + XS], gw[t]);
        }
      }
    }
  }
  if (dwdp) {                                     // CTA-uniform: reduce the K taps over the 256 threads through shared memory
    const int warp = threadIdx.x >> 5;
    __syncthreads();
#pragma unroll
    for (int t = 0; t < K; ++t) part[t * kThreads + threadIdx.x] = gw[t];
    __syncthreads();
    for (int v = warp; v < K; v += 8) {
      const float* pv = part + v * kThreads;
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < kThreads / 32; ++q) t += pv[lane + 32 * q];
      t = warp_sum(t);
      if (lane == 0) atomicAdd(&dwdp[v], t);
    }
  }
}

// CTA: channel c, 8 batch rows (one per warp), one segment of SEG input positions
template <int S>
__global__ void __launch_bounds__(kThreads, 3) k_edge_bwd(const __grid_constant__ EdgeBwdBatch P) {
  using G = EdgeBwdGeom<S>;
  constexpr int SEG = G::SEG, XO = G::XO, WX = G::WX, WD = G::WD;
  const EdgeBwdJob& J = P.j[blockIdx.z];
  const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rg = blockIdx.y / P.nseg, seg = blockIdx.y - rg * P.nseg;
  const int b = rg * 8 + warp;
  const int m0 = seg * SEG, m1 = min(J.Lin, m0 + SEG);
  const int l0 = m0 / S;
  const int l1 = m1 > m0 ? min(J.Lo, (m1 - 1) / S + 1) : l0;
  const bool ok = b < P.B && m0 < J.Lin;
  GNAS_DYN_SMEM(float, smem);
  float* xrow = smem + warp * G::ROW;      // relu(x) with zero halo
  float* drow = xrow + WX;                 // du of the current candidate with zero halo
  float* arow = drow + WD;                 // dX of the segment before masking
  float* part = smem + 8 * G::ROW;
  if (ok) {
    const float* xg = J.x + (long long)b * J.xbs + (long long)c * J.Lin;
    float xv_[(WX + 31) / 32];
#pragma unroll
    for (int q = 0; q < (WX + 31) / 32; ++q) {
      const int pos = m0 - XO + lane + 32 * q;
      xv_[q] = (pos >= 0 && pos < J.Lin) ? xg[pos] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < (WX + 31) / 32; ++q) {
      const int p = lane + 32 * q;
      if (p < WX) xrow[p] = xv_[q] > 0.f ? xv_[q] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < (SEG + 31) / 32; ++q) if (lane + 32 * q < SEG) arow[lane + 32 * q] = 0.f;
  }
  const long long ro = (long long)(ok ? b : 0) * J.dubs + (long long)c * J.Lo;
  if (J.du[0]) edge_bwd_cand<15, 1, S>(J.du[0] + ro, J.wd[0] + c * 15, J.dwd[0] ? J.dwd[0] + c * 15 : nullptr, J.Lo, xrow, drow, arow, part, m0, m1, l0, l1, ok);
  if (J.du[1]) edge_bwd_cand<9, 1, S>(J.du[1] + ro, J.wd[1] + c * 9, J.dwd[1] ? J.dwd[1] + c * 9 : nullptr, J.Lo, xrow, drow, arow, part, m0, m1, l0, l1, ok);
  if (J.du[2]) edge_bwd_cand<15, 2, S>(J.du[2] + ro, J.wd[2] + c * 15, J.dwd[2] ? J.dwd[2] + c * 15 : nullptr, J.Lo, xrow, drow, arow, part, m0, m1, l0, l1, ok);
  if (J.du[3]) edge_bwd_cand<9, 2, S>(J.du[3] + ro, J.wd[3] + c * 9, J.dwd[3] ? J.dwd[3] + c * 9 : nullptr, J.Lo, xrow, drow, arow, part, m0, m1, l0, l1, ok);
  __syncwarp();
  if (ok) {
    // one lane-interleaved read-modify-write of dX (full 128-byte lines whatever the row alignment)
    float* orow = J.out + (long long)b * J.obs + (long long)c * J.Lin;
    const int nm = m1 - m0;
    _Pragma("unroll 4")
    for (int ml = lane; ml < nm; ml += 32)
      if (xrow[ml + XO] > 0.f) orow[m0 + ml] += arow[ml];
  }
}

// ------------------------------------------------------------------ pools + identity skip backward
struct PoolBwdJob {
  const float* g; long long gbs;       // node gradient [B][C][Lo]
  const float* x; long long xbs;       // edge input [B][C][Lin]
  const float* zmax; const float* zavg;   // raw pool outputs (contiguous [B][C][Lo]) or null
  const float* coef_max; const float* coef_avg;
  const float* skip_mask;              // dropout mask or null
  float* dx; long long dxbs;           // accumulated
  int Lin, Lo, skip_widx;              // skip_widx < 0: no identity skip on this edge
};
struct PoolBwdBatch { int n, B, C, S, nseg; const float* mixw; PoolBwdJob j[kMaxPool]; };

// CTA: channel c, 8 batch rows (one per warp), one segment of kPoolSeg input positions (a multiple of every stride)
constexpr int kPoolSeg = 252, kPoolXO = 8, kPoolWX = kPoolSeg + 16, kPoolWD = kPoolSeg + 8;
static __global__ void __launch_bounds__(kThreads) k_pool_bwd(const __grid_constant__ PoolBwdBatch P) {
  const PoolBwdJob& J = P.j[blockIdx.z];
  const int c = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rg = blockIdx.y / P.nseg, seg = blockIdx.y - rg * P.nseg;
  const int b = rg * 8 + warp;
  const int S = P.S;
  const int m0 = seg * kPoolSeg, m1 = min(J.Lin, m0 + kPoolSeg);
  if (b >= P.B || m0 >= J.Lin) return;
  __shared__ float sm[8][kPoolWX + 2 * kPoolWD];
  __shared__ int smi[8][kPoolWD];
  float* xrow = sm[warp];        // index p <-> pos = m0 - kPoolXO + p
  float* dm = xrow + kPoolWX;    // dz of the max pool, index i <-> l = m0 / S - 4 + i
  float* da = dm + kPoolWD;      // dz of the avg pool already divided by the window count
  int* apos = smi[warp];         // position of the first maximal element of window l (torch max_pool1d backward routes there)
  const int lb = m0 / S - 4;
  const int nd = (m1 - m0 + 1) / S + 8;          // staged windows: l in [lb, lb + nd)
  const float* xg = J.x + (long long)b * J.xbs + (long long)c * J.Lin;
  const float* gg = J.g + (long long)b * J.gbs + (long long)c * J.Lo;
  const long long zo = ((long long)b * P.C + c) * J.Lo;
  float gv[(kPoolWD + 31) / 32], zm[(kPoolWD + 31) / 32], za[(kPoolWD + 31) / 32];
  {
    float xv[(kPoolWX + 31) / 32];
#pragma unroll
    for (int q = 0; q < (kPoolWX + 31) / 32; ++q) {
      const int pos = m0 - kPoolXO + lane + 32 * q;
      xv[q] = (J.zmax && pos >= 0 && pos < J.Lin) ? xg[pos] : -INFINITY;
    }
#pragma unroll
    for (int q = 0; q < (kPoolWD + 31) / 32; ++q) {
      const int i = lane + 32 * q, l = lb + i;
      const bool in = i < nd && l >= 0 && l < J.Lo;
      gv[q] = in ? gg[l] : 0.f;
      zm[q] = (in && J.zmax) ? J.zmax[zo + l] : 0.f;
      za[q] = (in && J.zavg) ? J.zavg[zo + l] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < (kPoolWX + 31) / 32; ++q) {
      const int p = lane + 32 * q;
      if (p < kPoolWX) xrow[p] = xv[q];
    }
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < (kPoolWD + 31) / 32; ++q) {
    const int i = lane + 32 * q, l = lb + i;
    if (i < kPoolWD) {
      float vm = 0.f, va = 0.f;
      int ap = -1;
      if (i < nd && l >= 0 && l < J.Lo) {
        if (J.zmax) {
          vm = fmaf(J.coef_max[c], gv[q], fmaf(J.coef_max[P.C + c], zm[q], J.coef_max[2 * P.C + c]));
          const int p0 = l * S - 2 - m0 + kPoolXO;           // xrow index of pos l*S-2; staged windows keep it in range
          if (p0 >= 0 && p0 + 4 < kPoolWX) {
            const float* wv = xrow + p0;
            int am = 0; float best = wv[0];
#pragma unroll
            for (int t = 1; t < 5; ++t) if (wv[t] > best) { best = wv[t]; am = t; }
            ap = l * S - 2 + am;
          }
        }
        if (J.zavg) {
          const int lo = max(l * S - 2, 0), hi = min(l * S + 2, J.Lin - 1);
          va = fmaf(J.coef_avg[c], gv[q], fmaf(J.coef_avg[P.C + c], za[q], J.coef_avg[2 * P.C + c])) / (float)(hi - lo + 1);
        }
      }
      dm[i] = vm; da[i] = va; apos[i] = ap;
    }
  }
  __syncwarp();
  const float wskip = J.skip_widx >= 0 ? P.mixw[J.skip_widx] : 0.f;
  float* drow = J.dx + (long long)b * J.dxbs + (long long)c * J.Lin;
  _Pragma("unroll 4")
  for (int m = m0 + lane; m < m1; m += 32) {
    float a = 0.f;
    // windows l with l*S-2 <= m <= l*S+2
    int l_lo = (m - 2 + S - 1) / S; if (m - 2 < 0) l_lo = 0;
    int l_hi = (m + 2) / S; if (l_hi > J.Lo - 1) l_hi = J.Lo - 1;
    for (int l = l_lo; l <= l_hi; ++l) {
      a += da[l - lb];
      if (apos[l - lb] == m) a += dm[l - lb];
    }
    if (J.skip_widx >= 0) {
      float gs = gg[m] * wskip;
      if (J.skip_mask) gs *= J.skip_mask[((long long)b * P.C + c) * J.Lin + m];
      a += gs;
    }
    drow[m] += a;
  }
}

// ------------------------------------------------------------------ small helpers
struct SliceCopy { const float* src; long long sbs; float* dst; long long dbs; int B, CL; };
static __global__ void __launch_bounds__(kThreads) k_copy_rows(const __grid_constant__ SliceCopy P) {
  const long long total = (long long)P.B * P.CL;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / P.CL);
    const int r = (int)(i - (long long)b * P.CL);
    P.dst[(long long)b * P.dbs + r] = P.src[(long long)b * P.sbs + r];
  }
}

static __global__ void k_double_to_float(const double* src, float* dst, int n, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = accumulate ? dst[i] + (float)src[i] : (float)src[i];
}

}  // namespace gnas

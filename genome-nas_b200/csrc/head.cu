// Classification head and the two loss terms of the search step, hand-written fp32.
//   flatten(permute(hiddens)) -> Dropout(0.1) -> Linear(T*H, 925) -> ReLU -> Linear(925, classes) -> Sigmoid
//       darts_tools/model_discCNN.py:547-566
//   BCELoss(mean) and the temporal activation regulariser beta * mean((h[1:] - h[:-1])^2)
//       generalNAS_tools/train_and_validate.py:136-140
// Work per pass at B = 64, T*H = 21504: three GEMMs of 2.5 GFLOP (forward, data gradient, weight gradient)
// around an 80 MB weight matrix -- each is read (or read-modified-written) exactly once per GEMM; the fp32 FFMA
// pipe, not HBM, bounds them (2.5 GFLOP / 72 TFLOP/s = 35 us against 80 MB / 6.5 TB/s = 12 us).
// One tiled kernel (64 x 64 x 16 tiles, 4 x 4 register tiles, register-staged double buffering, 128-bit loads
// wherever the operand allows) serves all five products through strides; split-K partials are reduced in a fixed
// order by the small epilogue kernels (bias, ReLU / Sigmoid, ReLU mask), so the head is deterministic.
#include "common.cuh"
#include "../../include/gnas.h"

namespace gnas {
namespace head {

constexpr int TM = 64, TN = 64, TK = 16, PITCH = 68;
enum : int { E_STORE = 0, E_ACC = 1, E_SCATTER = 2 };

struct GemmP {
  const float* A; long long sam, sak;     // A(m, k) = A[m * sam + k * sak]
  const float* B; long long sbk, sbn;     // B(k, n) = B[k * sbk + n * sbn]
  float* C; long long ldc, zstride;       // C(m, n) = C[z * zstride + m * ldc + n]
  int M, N, K, kchunk, epi;
  int a_kc, a_vec, b_kc, b_vec, c_vec;    // operand walk: K-contiguous (else row-contiguous), 128-bit accesses allowed
  const float* mask; int Hh, Bb;          // E_SCATTER: C is [T][Bb][Hh], n = t * Hh + h, value * mask[m][n]
};

// this thread's four elements of a 64 (rows) x 16 (k) operand tile
__device__ __forceinline__ void fetch(const float* X, long long sr, long long sk, int kc, int vec, int r0, int R,
                                      int k0, int kend, int tid, float (&v)[4]) {
  v[0] = v[1] = v[2] = v[3] = 0.f;
  if (kc) {
    const int r = r0 + (tid >> 2), k = k0 + (tid & 3) * 4;
    if (r < R && k < kend) {
      const float* p = X + (long long)r * sr + (long long)k * sk;
      if (vec && k + 3 < kend) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (k + i < kend) v[i] = p[(long long)i * sk];
      }
    }
  } else {
    const int k = k0 + (tid >> 4), r = r0 + (tid & 15) * 4;
    if (k < kend && r < R) {
      const float* p = X + (long long)k * sk + (long long)r * sr;
      if (vec && r + 3 < R) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) if (r + i < R) v[i] = p[(long long)i * sr];
      }
    }
  }
}
__device__ __forceinline__ void stash(float (*S)[PITCH], int kc, int tid, const float (&v)[4]) {
  if (kc) {
    const int r = tid >> 2, kq = (tid & 3) * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) S[kq + i][r] = v[i];
  } else {
    *reinterpret_cast<float4*>(&S[tid >> 4][(tid & 15) * 4]) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

static __global__ void __launch_bounds__(256) k_head_gemm(const __grid_constant__ GemmP P) {
  __shared__ __align__(16) float As[TK][PITCH];
  __shared__ __align__(16) float Bs[TK][PITCH];
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int kbeg = blockIdx.z * P.kchunk;
  const int kend = min(P.K, kbeg + P.kchunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[4], rb[4];
  fetch(P.A, P.sam, P.sak, P.a_kc, P.a_vec, m0, P.M, kbeg, kend, tid, ra);
  fetch(P.B, P.sbn, P.sbk, P.b_kc, P.b_vec, n0, P.N, kbeg, kend, tid, rb);
  for (int k0 = kbeg; k0 < kend; k0 += TK) {
    stash(As, P.a_kc, tid, ra);
    stash(Bs, P.b_kc, tid, rb);
    __syncthreads();
    if (k0 + TK < kend) {
      fetch(P.A, P.sam, P.sak, P.a_kc, P.a_vec, m0, P.M, k0 + TK, kend, tid, ra);
      fetch(P.B, P.sbn, P.sbk, P.b_kc, P.b_vec, n0, P.N, k0 + TK, kend, tid, rb);
    }
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int n = n0 + tx * 4;
  if (n >= P.N) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= P.M) continue;
    float v[4] = {acc[i][0], acc[i][1], acc[i][2], acc[i][3]};
    float* c;
    if (P.epi == E_SCATTER) {
      const int t = n / P.Hh, h = n - t * P.Hh;
      c = P.C + ((long long)t * P.Bb + m) * P.Hh + h;
      if (P.mask) {
        const float* mk = P.mask + (long long)m * P.N + n;
#pragma unroll
        for (int j = 0; j < 4; ++j) if (n + j < P.N) v[j] *= mk[j];
      }
    } else {
      c = P.C + (long long)blockIdx.z * P.zstride + (long long)m * P.ldc + n;
    }
    if (P.c_vec && n + 3 < P.N) {
      float4* c4 = reinterpret_cast<float4*>(c);
      if (P.epi == E_ACC) { const float4 o = *c4; v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w; }
      *c4 = make_float4(v[0], v[1], v[2], v[3]);
    } else if (P.epi == E_SCATTER) {
      // a group that straddles two time steps (H % 4 != 0): element-wise addressing
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < P.N) { const int t = (n + j) / P.Hh, h = (n + j) - t * P.Hh; P.C[((long long)t * P.Bb + m) * P.Hh + h] = v[j]; }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < P.N) c[j] = P.epi == E_ACC ? c[j] + v[j] : v[j];
    }
  }
}

// xd[b][t*H + h] = hs[t][b][h] * mask[b][t*H + h]   (permute(1,0,2) + flatten + dropout in one pass)
static __global__ void __launch_bounds__(256) k_head_gather(const float* __restrict__ hs, const float* __restrict__ mask,
                                                            float* __restrict__ xd, int T, int B, int H) {
  const long long F = (long long)T * H, total = (long long)B * F;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / F);
    const long long f = i - (long long)b * F;
    const int t = (int)(f / H), h = (int)(f - (long long)t * H);
    const float v = hs[((long long)t * B + b) * H + h];
    xd[i] = mask ? v * mask[i] : v;
  }
}

enum : int { R_RELU = 0, R_SIGMOID = 1, R_NONE = 2, R_RELU_MASK = 3 };
// out[m][n] = act(sum_z part[z][m][n] + bias[n]), fixed summation order; R_RELU_MASK: gate by ref[m][n] > 0
struct ReduceP { const float* part; long long zstride; int nz, M, N, ld, ldo, mode; const float* bias; const float* ref; float* out; };
static __global__ void __launch_bounds__(256) k_head_reduce(const __grid_constant__ ReduceP P) {
  const long long total = (long long)P.M * P.N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / P.N), n = (int)(i - (long long)m * P.N);
    const long long o = (long long)m * P.ld + n;
    float s = 0.f;
    for (int z = 0; z < P.nz; ++z) s += P.part[(long long)z * P.zstride + o];
    if (P.bias) s += P.bias[n];
    if (P.mode == R_RELU) s = s > 0.f ? s : 0.f;
    else if (P.mode == R_SIGMOID) s = 1.f / (1.f + expf(-s));
    else if (P.mode == R_RELU_MASK) s = P.ref[o] > 0.f ? s : 0.f;
    P.out[(long long)m * P.ldo + n] = s;
  }
}

// dz[m][n] = dout[m][n] * p (1 - p)   (Sigmoid backward; copy when the task has no final activation)
static __global__ void __launch_bounds__(256) k_head_dact(const float* __restrict__ dout, const float* __restrict__ p,
                                                          float* __restrict__ dz, int M, int N, int ld, int sigmoid) {
  const long long total = (long long)M * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(i / N), n = (int)(i - (long long)m * N);
    const float g = dout[i];
    float v = g;
    if (sigmoid) { const float q = p[i]; v = g * q * (1.f - q); }
    dz[(long long)m * ld + n] = v;
  }
}

// dbias[n] += sum_m X[m][n]
static __global__ void __launch_bounds__(256) k_head_colsum(const float* __restrict__ X, int M, int N, int ld, float* __restrict__ db) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int m = 0; m < M; ++m) s += X[(long long)m * ld + n];
  db[n] += s;
}

// ---- losses: partial sums in double per CTA, then one CTA adds them in index order (deterministic)
constexpr int LOSS_CTAS = 148;
// BCELoss term, torch semantics: log clamped at -100
static __global__ void __launch_bounds__(256) k_bce_partial(const float* __restrict__ p, const float* __restrict__ y, long long n, double* part) {
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float q = p[i], t = y[i];
    const float lp = fmaxf(logf(q), -100.f), lq = fmaxf(log1pf(-q), -100.f);
    s -= (double)(t * lp + (1.f - t) * lq);
  }
  s = warp_sum_d(s);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < 8; ++i) t += red[i]; part[blockIdx.x] = t; }
}
// sum over t < T-1 of (h[t+1] - h[t])^2, h = [T][BH]
static __global__ void __launch_bounds__(256) k_tar_partial(const float* __restrict__ h, long long n, long long BH, double* part) {
  double s = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = h[i + BH] - h[i];
    s += (double)d * (double)d;
  }
  s = warp_sum_d(s);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int i = 0; i < 8; ++i) t += red[i]; part[blockIdx.x] = t; }
}
static __global__ void k_loss_final(const double* part, int n, double inv, float* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < n; ++i) t += part[i];
    *out = (float)(t * inv);
  }
}
// dp = gout * (p - y) / max(p (1 - p), 1e-12) / n   (torch binary_cross_entropy_backward)
static __global__ void __launch_bounds__(256) k_bce_bwd(const float* __restrict__ p, const float* __restrict__ y, const float* gout,
                                                        long long n, float* __restrict__ dp) {
  const float g = *gout / (float)n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float q = p[i];
    dp[i] = g * (q - y[i]) / fmaxf(q * (1.f - q), 1e-12f);
  }
}
// dh[t] = gout * 2 / n * ((h[t] - h[t-1]) [t > 0] - (h[t+1] - h[t]) [t < T-1])
static __global__ void __launch_bounds__(256) k_tar_bwd(const float* __restrict__ h, const float* gout, int T, long long BH,
                                                        float* __restrict__ dh) {
  const long long total = (long long)T * BH;
  const float g = *gout * 2.f / (float)((long long)(T - 1) * BH);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / BH);
    const float c = h[i];
    float d = 0.f;
    if (t > 0) d += c - h[i - BH];
    if (t < T - 1) d -= h[i + BH] - c;
    dh[i] = g * d;
  }
}

// ---------------------------------------------------------------- host side
struct Plan {
  int F, D1p, NCp, S1, k1, S2, k2, S3, k3;
  long long xd, hid, part1, part2, fwd_total;      // forward workspace (floats); xd and hid are read by backward
  long long dz, part3, dhp, bwd_total;             // backward scratch
};
static inline int chunk_of(int K, int want, int max_splits, int* splits) {
  int s = cdiv(K, want);
  if (s > max_splits) s = max_splits;
  if (s < 1) s = 1;
  const int k = round_up(cdiv(K, s), TK);
  *splits = cdiv(K, k);
  return k;
}
static Plan plan(int T, int B, int H, int D1, int NC) {
  Plan p;
  p.F = T * H; p.D1p = round_up(D1, 4); p.NCp = round_up(NC, 4);
  p.k1 = chunk_of(p.F, 512, 64, &p.S1);
  p.k2 = chunk_of(D1, 192, 8, &p.S2);
  p.k3 = chunk_of(NC, 192, 8, &p.S3);
  long long o = 0;
  p.xd = o; o += (long long)B * p.F; o = (o + 3) / 4 * 4;
  p.hid = o; o += (long long)B * p.D1p;
  p.part1 = o; o += (long long)p.S1 * B * p.D1p;
  p.part2 = o; o += (long long)p.S2 * B * p.NCp;
  p.fwd_total = o;
  o = 0;
  p.dz = o; o += (long long)B * p.NCp;
  p.part3 = o; o += (long long)p.S3 * B * p.D1p;
  p.dhp = o; o += (long long)B * p.D1p;
  p.bwd_total = o;
  return p;
}
static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static void gemm(cudaStream_t st, const float* A, long long sam, long long sak, const float* Bm, long long sbk, long long sbn,
                 float* Cm, long long ldc, long long zstride, int M, int N, int K, int kchunk, int splits, int epi,
                 const float* mask = nullptr, int Hh = 1, int Bb = 1) {
  GemmP P;
  P.A = A; P.sam = sam; P.sak = sak; P.B = Bm; P.sbk = sbk; P.sbn = sbn; P.C = Cm; P.ldc = ldc; P.zstride = zstride;
  P.M = M; P.N = N; P.K = K; P.kchunk = kchunk; P.epi = epi; P.mask = mask; P.Hh = Hh; P.Bb = Bb;
  P.a_kc = (sak == 1 || sam != 1) ? 1 : 0;
  P.a_vec = al16(A) && (P.a_kc ? (sak == 1 && sam % 4 == 0) : (sam == 1 && sak % 4 == 0));
  P.b_kc = (sbk == 1 || sbn != 1) ? 1 : 0;
  P.b_vec = al16(Bm) && (P.b_kc ? (sbk == 1 && sbn % 4 == 0) : (sbn == 1 && sbk % 4 == 0));
  if (epi == E_SCATTER) P.c_vec = al16(Cm) && Hh % 4 == 0 && (!mask || (al16(mask) && N % 4 == 0));
  else P.c_vec = al16(Cm) && ldc % 4 == 0 && zstride % 4 == 0;
  GNAS_LAUNCH(k_head_gemm, dim3(cdiv(N, TN), cdiv(M, TM), splits), dim3(256), 0, st, P);
}
static inline int blocks_for(long long n) {
  long long b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}
static void reduce(cudaStream_t st, const float* part, long long zstride, int nz, int M, int N, int ld, int ldo, int mode,
                   const float* bias, const float* ref, float* out) {
  ReduceP R;
  R.part = part; R.zstride = zstride; R.nz = nz; R.M = M; R.N = N; R.ld = ld; R.ldo = ldo; R.mode = mode; R.bias = bias; R.ref = ref; R.out = out;
  GNAS_LAUNCH(k_head_reduce, dim3(blocks_for((long long)M * N)), dim3(256), 0, st, R);
}

}  // namespace head
}  // namespace gnas

using namespace gnas;
using namespace gnas::head;

static bool head_args_ok(int T, int B, int H, int D1, int NC) { return T >= 1 && B >= 1 && H >= 1 && D1 >= 1 && NC >= 1; }

extern "C" int gnas_head_workspace(int T, int B, int H, int D1, int NC, int64_t* fwd_floats, int64_t* bwd_floats) {
  if (!head_args_ok(T, B, H, D1, NC) || !fwd_floats || !bwd_floats) { set_error("head_workspace: bad argument"); return GNAS_ERR_ARG; }
  const Plan p = plan(T, B, H, D1, NC);
  *fwd_floats = p.fwd_total;
  *bwd_floats = p.bwd_total;
  return GNAS_OK;
}

extern "C" int gnas_head_forward(int T, int B, int H, int D1, int NC, int sigmoid, const float* hs, const float* drop_mask,
                                 const float* Wd, const float* bd, const float* Wc, const float* bc, float* out,
                                 float* ws, void* stream) {
  if (!head_args_ok(T, B, H, D1, NC) || !hs || !Wd || !Wc || !out || !ws) { set_error("head_forward: bad argument"); return GNAS_ERR_ARG; }
  if (!al16(ws)) { set_error("head_forward: workspace must be 16-byte aligned"); return GNAS_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const Plan p = plan(T, B, H, D1, NC);
  float* xd = ws + p.xd; float* hid = ws + p.hid; float* part1 = ws + p.part1; float* part2 = ws + p.part2;
  GNAS_LAUNCH(k_head_gather, dim3(blocks_for((long long)B * p.F)), dim3(256), 0, st, hs, drop_mask, xd, T, B, H);
  // hid[b][o] = relu(sum_f xd[b][f] Wd[o][f] + bd[o])
  gemm(st, xd, p.F, 1, Wd, 1, p.F, part1, p.D1p, (long long)B * p.D1p, B, D1, p.F, p.k1, p.S1, E_STORE);
  reduce(st, part1, (long long)B * p.D1p, p.S1, B, D1, p.D1p, p.D1p, R_RELU, bd, nullptr, hid);
  // out[b][j] = sigmoid(sum_o hid[b][o] Wc[j][o] + bc[j])
  gemm(st, hid, p.D1p, 1, Wc, 1, D1, part2, p.NCp, (long long)B * p.NCp, B, NC, D1, p.k2, p.S2, E_STORE);
  reduce(st, part2, (long long)B * p.NCp, p.S2, B, NC, p.NCp, NC, sigmoid ? R_SIGMOID : R_NONE, bc, nullptr, out);
  return check_launch("gnas_head_forward");
}

/* dhs is overwritten; dWd / dbd / dWc / dbc are accumulated (null or need_wgrad = 0: skipped) */
extern "C" int gnas_head_backward(int T, int B, int H, int D1, int NC, int sigmoid, const float* drop_mask, const float* Wd,
                                  const float* Wc, const float* out, const float* dout, const float* ws, float* scratch,
                                  float* dhs, float* dWd, float* dbd, float* dWc, float* dbc, int need_wgrad, void* stream) {
  if (!head_args_ok(T, B, H, D1, NC) || !Wd || !Wc || !out || !dout || !ws || !scratch || !dhs) {
    set_error("head_backward: bad argument"); return GNAS_ERR_ARG;
  }
  if (need_wgrad && (!dWd || !dWc)) { set_error("head_backward: weight gradients requested without buffers"); return GNAS_ERR_ARG; }
  if (!al16(ws) || !al16(scratch)) { set_error("head_backward: workspaces must be 16-byte aligned"); return GNAS_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const Plan p = plan(T, B, H, D1, NC);
  const float* xd = ws + p.xd; const float* hid = ws + p.hid;
  float* dz = scratch + p.dz; float* part3 = scratch + p.part3; float* dhp = scratch + p.dhp;
  GNAS_LAUNCH(k_head_dact, dim3(blocks_for((long long)B * NC)), dim3(256), 0, st, dout, out, dz, B, NC, p.NCp, sigmoid);
  if (need_wgrad) {
    // dWc[j][o] += sum_b dz[b][j] hid[b][o];  dbc[j] += sum_b dz[b][j]
    gemm(st, dz, 1, p.NCp, hid, p.D1p, 1, dWc, D1, 0, NC, D1, B, round_up(B, TK), 1, E_ACC);
    if (dbc) GNAS_LAUNCH(k_head_colsum, dim3(cdiv(NC, 256)), dim3(256), 0, st, (const float*)dz, B, NC, p.NCp, dbc);
  }
  // dhp[b][o] = (hid > 0) sum_j dz[b][j] Wc[j][o]
  gemm(st, dz, p.NCp, 1, Wc, D1, 1, part3, p.D1p, (long long)B * p.D1p, B, D1, NC, p.k3, p.S3, E_STORE);
  reduce(st, part3, (long long)B * p.D1p, p.S3, B, D1, p.D1p, p.D1p, R_RELU_MASK, nullptr, hid, dhp);
  if (need_wgrad) {
    // dWd[o][f] += sum_b dhp[b][o] xd[b][f];  dbd[o] += sum_b dhp[b][o]
    gemm(st, dhp, 1, p.D1p, xd, p.F, 1, dWd, p.F, 0, D1, p.F, B, round_up(B, TK), 1, E_ACC);
    if (dbd) GNAS_LAUNCH(k_head_colsum, dim3(cdiv(D1, 256)), dim3(256), 0, st, (const float*)dhp, B, D1, p.D1p, dbd);
  }
  // dhs[t][b][h] = mask[b][f] sum_o dhp[b][o] Wd[o][f],  f = t H + h
  gemm(st, dhp, p.D1p, 1, Wd, p.F, 1, dhs, 0, 0, B, p.F, D1, round_up(D1, TK), 1, E_SCATTER, drop_mask, H, B);
  return check_launch("gnas_head_backward");
}

/* ws: LOSS_CTAS doubles (gnas_loss_workspace floats) */
extern "C" int gnas_loss_workspace(int64_t* floats) {
  if (!floats) { set_error("loss_workspace: bad argument"); return GNAS_ERR_ARG; }
  *floats = 2 * LOSS_CTAS + 2;
  return GNAS_OK;
}
static inline double* dbl_ws(float* ws) { return reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(ws) + 7) & ~(uintptr_t)7); }

extern "C" int gnas_bce_forward(const float* p, const float* y, int64_t n, float* loss, float* ws, void* stream) {
  if (!p || !y || !loss || !ws || n < 1) { set_error("bce_forward: bad argument"); return GNAS_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int ctas = (int)(cdivll(n, 2048) < LOSS_CTAS ? cdivll(n, 2048) : LOSS_CTAS);
  GNAS_LAUNCH(k_bce_partial, dim3(ctas), dim3(256), 0, st, p, y, (long long)n, dbl_ws(ws));
  GNAS_LAUNCH(k_loss_final, dim3(1), dim3(32), 0, st, (const double*)dbl_ws(ws), ctas, 1.0 / (double)n, loss);
  return check_launch("gnas_bce_forward");
}
extern "C" int gnas_bce_backward(const float* p, const float* y, const float* gout, int64_t n, float* dp, void* stream) {
  if (!p || !y || !gout || !dp || n < 1) { set_error("bce_backward: bad argument"); return GNAS_ERR_ARG; }
  GNAS_LAUNCH(k_bce_bwd, dim3(blocks_for(n)), dim3(256), 0, (cudaStream_t)stream, p, y, gout, (long long)n, dp);
  return check_launch("gnas_bce_backward");
}
extern "C" int gnas_tar_forward(const float* h, int T, int64_t BH, float* loss, float* ws, void* stream) {
  if (!h || !loss || !ws || T < 2 || BH < 1) { set_error("tar_forward: bad argument"); return GNAS_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = (long long)(T - 1) * BH;
  const int ctas = (int)(cdivll(n, 2048) < LOSS_CTAS ? cdivll(n, 2048) : LOSS_CTAS);
  GNAS_LAUNCH(k_tar_partial, dim3(ctas), dim3(256), 0, st, h, n, (long long)BH, dbl_ws(ws));
  GNAS_LAUNCH(k_loss_final, dim3(1), dim3(32), 0, st, (const double*)dbl_ws(ws), ctas, 1.0 / (double)n, loss);
  return check_launch("gnas_tar_forward");
}
extern "C" int gnas_tar_backward(const float* h, const float* gout, int T, int64_t BH, float* dh, void* stream) {
  if (!h || !gout || !dh || T < 2 || BH < 1) { set_error("tar_backward: bad argument"); return GNAS_ERR_ARG; }
  GNAS_LAUNCH(k_tar_bwd, dim3(blocks_for((long long)T * BH)), dim3(256), 0, (cudaStream_t)stream, h, gout, T, (long long)BH, dh);
  return check_launch("gnas_tar_backward");
}

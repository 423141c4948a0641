// Host-side launch helpers: pick tile shapes, size shared memory, dispatch the
// template instance.  One process per GPU; everything is enqueued on the
// caller's stream and nothing here synchronises.
#pragma once
#include "cnn_fwd.cuh"
#include "cnn_bwd.cuh"

namespace gnas {

#ifdef GNAS_EMUL
#define GNAS_PREP_SMEM(kernel, bytes) do { } while (0)
#else
#define GNAS_PREP_SMEM(kernel, bytes)                                                       \
  do {                                                                                      \
    static int prepared_bytes[kMaxDevices] = {0};                                           \
    if (need_prepare(prepared_bytes, (int)(bytes)))                                         \
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)); \
  } while (0)
#endif

static inline int choose_tm(int C) {
  if (C % 8 == 0 && C >= 32) return 8;
  if (C % 4 == 0 && C >= 16) return 4;
  return 2;
}
static inline int gcd_int(int a, int b) { while (b) { const int t = a % b; a = b; b = t; } return a; }
// channels of the reduction dimension staged per shared-memory chunk.  Pointwise (KW = 1) weights are small:
// take as many channels as fit ~96 KB so that a CTA loads once and computes once; KW >= 13 streams 8 at a time.
static inline int choose_kc(int kw, const int* cins, int n, int cout, int wp) {
  int g = cins[0];
  for (int i = 1; i < n; ++i) g = gcd_int(g, cins[i]);
  if (kw != 1) { int kc = cout > 64 ? 4 : 8; while (g % kc) kc >>= 1; return kc; }   // <= 32 KB of weights per chunk
  int kc = g;
  while (kc > 8 && kc % 2 == 0 && sizeof(float) * ((size_t)kc * wp + (size_t)kc * cout) > 48 * 1024) kc /= 2;
  return kc;
}

// ------------------------------------------------------------------ dense forward
template <int KW, int S, int TM>
static void launch_dense_fwd_t(DenseBatch& b, int maxLout, cudaStream_t st) {
  const int ncog = b.Cout / TM;
  const int nlg = kThreads / ncog;
  b.TL = 4 * nlg;
  const int Wp = ((b.TL - 1) * S + KW + 3) / 4 * 4 + 4;
  int cins[kMaxDense];
  for (int i = 0; i < b.n; ++i) cins[i] = b.j[i].Cin;
  b.KC = choose_kc(KW, cins, b.n, b.Cout, Wp);
  const size_t smem = sizeof(float) * (2 * ((size_t)b.KC * Wp + (size_t)b.KC * KW * b.Cout) + 4 * b.Cout);
  GNAS_PREP_SMEM((k_dense_fwd<KW, S, TM>), smem);
  dim3 grid(cdiv(maxLout, b.TL), b.B, b.n);
  GNAS_LAUNCH((k_dense_fwd<KW, S, TM>), grid, dim3(kThreads), smem, st, b);
}
template <int KW, int S>
static void launch_dense_fwd_s(DenseBatch& b, int maxLout, cudaStream_t st) {
  int tm = choose_tm(b.Cout);
  // 1x1 convolutions (preprocess, FactorizedReduce) have no arithmetic to amortise: narrower register tiles give
  // shorter position tiles and hence more CTAs (TM = 8 left cells 2-5 with 64 CTAs of ~90 us on 148 SMs)
  if (KW == 1)
    while (tm > 2 && (long long)cdiv(maxLout, 4 * (kThreads / (b.Cout / tm))) * b.B * b.n < 2 * 148) tm >>= 1;
  switch (tm) {
    case 8: launch_dense_fwd_t<KW, S, 8>(b, maxLout, st); break;
    case 4: launch_dense_fwd_t<KW, S, 4>(b, maxLout, st); break;
    default: launch_dense_fwd_t<KW, S, 2>(b, maxLout, st); break;
  }
}
// KW in {1, 13, 15}, S in {1, 2, 3}
static int launch_dense_fwd(int KW, int S, DenseBatch& b, cudaStream_t st) {
  if (b.n == 0) return 0;
  int maxL = 0;
  for (int i = 0; i < b.n; ++i) maxL = b.j[i].Lout > maxL ? b.j[i].Lout : maxL;
  if (KW == 15 && S == 1) launch_dense_fwd_s<15, 1>(b, maxL, st);
  else if (KW == 15 && S == 2) launch_dense_fwd_s<15, 2>(b, maxL, st);
  else if (KW == 15 && S == 3) launch_dense_fwd_s<15, 3>(b, maxL, st);
  else if (KW == 1 && S == 1) launch_dense_fwd_s<1, 1>(b, maxL, st);
  else if (KW == 1 && S == 2) launch_dense_fwd_s<1, 2>(b, maxL, st);
  else if (KW == 1 && S == 3) launch_dense_fwd_s<1, 3>(b, maxL, st);
  else if (KW == 13 && S == 1) launch_dense_fwd_s<13, 1>(b, maxL, st);
  else { set_error("dense_fwd: unsupported (KW, stride)"); return GNAS_ERR_UNSUPPORTED; }
  b.n = 0;
  return 0;
}

// ------------------------------------------------------------------ transposed dense (dx)
template <int KW, int S, int TM>
static void launch_dense_dx_t(DxBatch& b, int maxLin, cudaStream_t st) {
  const int ncig = b.Cin / TM;
  b.nq = kThreads / (ncig * S);
  constexpr int NT = (KW + S - 1) / S;
  const int Wu = 4 * b.nq + NT + 1;
  const int Wp = (Wu + 3) / 4 * 4 + 4;
  int couts[kMaxDense];
  for (int i = 0; i < b.n; ++i) couts[i] = b.j[i].Cout;
  b.KC = choose_kc(KW, couts, b.n, b.Cin, Wp);
  const size_t smem = sizeof(float) * (2 * ((size_t)b.KC * Wp + (size_t)b.KC * KW * b.Cin) + 4 * b.Cin);
  GNAS_PREP_SMEM((k_dense_bwd_dx<KW, S, TM>), smem);
  dim3 grid(cdiv(maxLin, 4 * S * b.nq), b.B, b.n);
  GNAS_LAUNCH((k_dense_bwd_dx<KW, S, TM>), grid, dim3(kThreads), smem, st, b);
}
template <int KW, int S>
static void launch_dense_dx_s(DxBatch& b, int maxLin, cudaStream_t st) {
  switch (choose_tm(b.Cin)) {
    case 8: launch_dense_dx_t<KW, S, 8>(b, maxLin, st); break;
    case 4: launch_dense_dx_t<KW, S, 4>(b, maxLin, st); break;
    default: launch_dense_dx_t<KW, S, 2>(b, maxLin, st); break;
  }
}
static int launch_dense_dx(int KW, int S, DxBatch& b, cudaStream_t st) {
  if (b.n == 0) return 0;
  int maxL = 0;
  for (int i = 0; i < b.n; ++i) maxL = b.j[i].Lin > maxL ? b.j[i].Lin : maxL;
  if (KW == 15 && S == 1) launch_dense_dx_s<15, 1>(b, maxL, st);
  else if (KW == 15 && S == 2) launch_dense_dx_s<15, 2>(b, maxL, st);
  else if (KW == 15 && S == 3) launch_dense_dx_s<15, 3>(b, maxL, st);
  else if (KW == 1 && S == 1) launch_dense_dx_s<1, 1>(b, maxL, st);
  else if (KW == 1 && S == 2) launch_dense_dx_s<1, 2>(b, maxL, st);
  else if (KW == 1 && S == 3) launch_dense_dx_s<1, 3>(b, maxL, st);
  else { set_error("dense_dx: unsupported (KW, stride)"); return GNAS_ERR_UNSUPPORTED; }
  b.n = 0;
  return 0;
}

// ------------------------------------------------------------------ weight gradients
static inline int pick_tc(int a, int b) {
  const int m = a < b ? a : b;
  if (a % 8 == 0 && b % 8 == 0 && m >= 64) return 8;
  if (a % 4 == 0 && b % 4 == 0 && m >= 16) return 4;
  if (a % 2 == 0 && b % 2 == 0 && m >= 8) return 2;
  return 1;
}
// CTAs per launch to aim for when splitting the position range of a weight-gradient job
constexpr int kWgradTargetCtas = 2 * 148;

template <int TC>
static void launch_pw_wgrad_t(WgBatch& b, int nblk, int ny, cudaStream_t st) {
  const size_t smem = sizeof(float) * (size_t)b.TL * (b.COB + 4 + b.CIB + 4);
  GNAS_PREP_SMEM((k_pw_wgrad<TC>), smem);
  dim3 grid(nblk, ny, b.n);
  GNAS_LAUNCH((k_pw_wgrad<TC>), grid, dim3(kThreads), smem, st, b);
}
// all jobs of a batch must share Cin and Lo (true for every call site)
static int launch_pw_wgrad(int S, WgBatch& b, cudaStream_t st) {
  if (b.n == 0) return 0;
  const int Cin = b.j[0].Cin, Lo = b.j[0].Lo;
  for (int i = 1; i < b.n; ++i) if (b.j[i].Cin != Cin || b.j[i].Lo != Lo) { set_error("pw_wgrad: mixed shapes"); return GNAS_ERR_ARG; }
  int TC = pick_tc(b.Cout, Cin);
  b.S = S; b.pad = 0;
  int nblk = 1, ny = 1, tpc = 0;
  auto configure = [&](int tc) -> bool {
    b.COB = b.Cout < 16 * tc ? b.Cout : 16 * tc;
    b.CIB = Cin < 16 * tc ? Cin : 16 * tc;
    while ((b.COB / tc) * (b.CIB / tc) > kThreads) b.CIB /= 2;
    if (b.Cout % b.COB || Cin % b.CIB || b.COB % tc || b.CIB % tc) return false;
    const int cmax = b.COB > b.CIB ? b.COB : b.CIB;
    int tl = 4096 / cmax;
    tl = tl > 512 ? 512 : (tl < 64 ? 64 : tl);
    b.TL = tl;
    nblk = (b.Cout / b.COB) * (Cin / b.CIB);
    const int tiles = b.B * cdiv(Lo, b.TL);
    tpc = (int)((long long)tiles * nblk * b.n / kWgradTargetCtas);
    b.TPC = tpc < 1 ? 1 : tpc;
    ny = cdiv(tiles, b.TPC);
    return true;
  };
  if (!configure(TC)) { set_error("pw_wgrad: channel counts must tile"); return GNAS_ERR_SHAPE; }
  // Every CTA flushes its COB x CIB block with atomics once: with few jobs and wide blocks the position range is
  // cut into many chunks and the flushes (ny * Cout * Cin atomics on the same addresses) dominate.  Smaller
  // register tiles give more blocks, hence fewer, longer chunks.
  while (tpc < 4 && TC > 1) {
    if (!configure(TC / 2)) { configure(TC); break; }
    TC /= 2;
  }
  switch (TC) {
    case 8: launch_pw_wgrad_t<8>(b, nblk, ny, st); break;
    case 4: launch_pw_wgrad_t<4>(b, nblk, ny, st); break;
    case 2: launch_pw_wgrad_t<2>(b, nblk, ny, st); break;
    default: launch_pw_wgrad_t<1>(b, nblk, ny, st); break;
  }
  b.n = 0;
  return 0;
}

template <int KW, int S>
static void launch_conv_wgrad_t(WgBatch& b, int nblk, int Lo, cudaStream_t st) {
  const int per = (b.COB / 4) * b.CIB;
  const int G = kThreads / per;
  int tlg = 512 / G;
  tlg = tlg > 64 ? 64 : (tlg < 16 ? 16 : tlg);
  while (tlg > 16 && tlg * G >= 2 * Lo) tlg /= 2;     // short rows: do not pad a tile to twice the row
  tlg &= ~3;                                          // groups start on 16-byte boundaries of the dz rows
  b.TL = tlg;
  const int TLc = tlg * G;
  // r rows are read up to 3*S past the last used window element (4-position steps): keep a small tail
  const int pz = round_up(TLc, 4) + 4, pr = ((TLc - 1) * S + KW) | 1;
  const size_t smem = sizeof(float) * ((size_t)b.COB * pz + (size_t)b.CIB * pr + 4 * S + 8);
  GNAS_PREP_SMEM((k_conv_wgrad<KW, S>), smem);
  const int tiles = b.B * cdiv(Lo, TLc);
  int tpc = (int)((long long)tiles * nblk * b.n / kWgradTargetCtas);
  b.TPC = tpc < 1 ? 1 : tpc;
  dim3 grid(nblk, cdiv(tiles, b.TPC), b.n);
  GNAS_LAUNCH((k_conv_wgrad<KW, S>), grid, dim3(kThreads), smem, st, b);
}
static int launch_conv_wgrad(int KW, int S, int pad, WgBatch& b, cudaStream_t st) {
  if (b.n == 0) return 0;
  const int Cin = b.j[0].Cin, Lo = b.j[0].Lo;
  for (int i = 1; i < b.n; ++i) if (b.j[i].Cin != Cin || b.j[i].Lo != Lo) { set_error("conv_wgrad: mixed shapes"); return GNAS_ERR_ARG; }
  b.S = S; b.pad = pad;
  b.COB = b.Cout < 32 ? b.Cout : 32;
  b.CIB = Cin < 32 ? Cin : 32;
  if (b.Cout % b.COB || Cin % b.CIB || b.COB % 4) { set_error("conv_wgrad: channel counts must tile"); return GNAS_ERR_SHAPE; }
  const int nblk = (b.Cout / b.COB) * (Cin / b.CIB);
  if (KW == 15 && S == 1) launch_conv_wgrad_t<15, 1>(b, nblk, Lo, st);
  else if (KW == 15 && S == 2) launch_conv_wgrad_t<15, 2>(b, nblk, Lo, st);
  else if (KW == 15 && S == 3) launch_conv_wgrad_t<15, 3>(b, nblk, Lo, st);
  else if (KW == 13 && S == 1) launch_conv_wgrad_t<13, 1>(b, nblk, Lo, st);
  else { set_error("conv_wgrad: unsupported (KW, stride)"); return GNAS_ERR_UNSUPPORTED; }
  b.n = 0;
  return 0;
}

// ------------------------------------------------------------------ depthwise (second stage of sep_conv)
// second depthwise convolution of the sep_conv candidates: K = 15 and K = 9 jobs in one launch
static int launch_dw2_fwd(DwBatch& b, cudaStream_t st) {
  if (b.n == 0) return 0;
  int lout = 1;
  for (int i = 0; i < b.n; ++i) lout = b.j[i].Lout > lout ? b.j[i].Lout : lout;
  b.nseg = cdiv(lout, DwFwdGeom<15, 1, 1>::SEGO);
  dim3 grid(cdiv(b.C, 8), b.B * b.nseg, b.n);
  GNAS_LAUNCH(k_dw2_fwd, grid, dim3(kThreads), 0, st, b);
  b.n = 0;
  return 0;
}
static int launch_dw2_bwd(DwBwdBatch& b, cudaStream_t st) {
  if (b.n == 0) return 0;
  using G = DwBwdGeom<15, 1, 1>;
  int lin = 1;
  for (int i = 0; i < b.n; ++i) lin = b.j[i].Lin > lin ? b.j[i].Lin : lin;
  b.nseg = cdiv(lin, G::SEG);
  const size_t smem = sizeof(float) * 8 * (size_t)(G::WX + G::WD + G::SEG);
  GNAS_PREP_SMEM(k_dw2_bwd, smem);
  int rpw = (b.C * cdiv(b.B, 8) * b.n * b.nseg) / (16 * 148);
  rpw = rpw < 1 ? 1 : (rpw > 2 ? 2 : rpw);
  b.RPW = rpw;
  dim3 grid(b.C, cdiv(b.B, 8 * rpw) * b.nseg, b.n);
  GNAS_LAUNCH(k_dw2_bwd, grid, dim3(kThreads), smem, st, b);
  b.n = 0;
  return 0;
}

// ------------------------------------------------------------------ fused edge forward
template <int S>
static void launch_edge_fwd_t(EdgeFwdBatch& b, cudaStream_t st) {
  int lout = 1;
  for (int i = 0; i < b.n; ++i) lout = b.j[i].Lout > lout ? b.j[i].Lout : lout;
  b.nseg = cdiv(lout, EdgeFwdGeom<S>::SEGO);
  dim3 grid(cdiv(b.C, 8), b.B * b.nseg, b.n);
  GNAS_LAUNCH((k_edge_fwd<S>), grid, dim3(kThreads), 0, st, b);
}
static int launch_edge_fwd(int S, EdgeFwdBatch& b, cudaStream_t st) {
  if (b.n == 0) return 0;
  if (S == 1) launch_edge_fwd_t<1>(b, st);
  else if (S == 2) launch_edge_fwd_t<2>(b, st);
  else if (S == 3) launch_edge_fwd_t<3>(b, st);
  else { set_error("edge_fwd: unsupported stride"); return GNAS_ERR_UNSUPPORTED; }
  b.n = 0;
  return 0;
}

// ------------------------------------------------------------------ fused edge backward
template <int S>
static void launch_edge_bwd_t(EdgeBwdBatch& b, cudaStream_t st) {
  using G = EdgeBwdGeom<S>;
  int lin = 1;
  for (int i = 0; i < b.n; ++i) lin = b.j[i].Lin > lin ? b.j[i].Lin : lin;
  b.nseg = cdiv(lin, G::SEG);
  const size_t smem = sizeof(float) * (8 * (size_t)G::ROW + G::PART);
  GNAS_PREP_SMEM((k_edge_bwd<S>), smem);
  dim3 grid(b.C, cdiv(b.B, 8) * b.nseg, b.n);
  GNAS_LAUNCH((k_edge_bwd<S>), grid, dim3(kThreads), smem, st, b);
}
static int launch_edge_bwd(int S, EdgeBwdBatch& b, cudaStream_t st) {
  if (b.n == 0) return 0;
  if (S == 1) launch_edge_bwd_t<1>(b, st);
  else if (S == 2) launch_edge_bwd_t<2>(b, st);
  else if (S == 3) launch_edge_bwd_t<3>(b, st);
  else { set_error("edge_bwd: unsupported stride"); return GNAS_ERR_UNSUPPORTED; }
  b.n = 0;
  return 0;
}

// ------------------------------------------------------------------ pools
static int launch_pool_fwd(PoolBatch& b, cudaStream_t st) {
  if (b.n == 0) return 0;
  int lout = 1;
  for (int i = 0; i < b.n; ++i) lout = b.j[i].Lout > lout ? b.j[i].Lout : lout;
  b.nseg = cdiv(lout, kPoolFwdSeg);
  dim3 grid(b.C, cdiv(b.B, 8) * b.nseg, b.n);
  GNAS_LAUNCH(k_pool_fwd, grid, dim3(kThreads), 0, st, b);
  b.n = 0;
  return 0;
}
static int launch_pool_bwd(PoolBwdBatch& b, cudaStream_t st) {
  if (b.n == 0) return 0;
  int lin = 1;
  for (int i = 0; i < b.n; ++i) lin = b.j[i].Lin > lin ? b.j[i].Lin : lin;
  b.nseg = cdiv(lin, kPoolSeg);
  dim3 grid(b.C, cdiv(b.B, 8) * b.nseg, b.n);
  GNAS_LAUNCH(k_pool_bwd, grid, dim3(kThreads), 0, st, b);
  b.n = 0;
  return 0;
}

static inline int ew_blocks(long long total) {
  long long nb = cdivll(total, kThreads);
  return (int)(nb < 1 ? 1 : (nb > 148 * 16 ? 148 * 16 : nb));
}

}  // namespace gnas

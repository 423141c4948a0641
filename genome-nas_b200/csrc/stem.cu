// Stem: Conv1d(4, C0, 13, padding 6, bias=False) + BatchNorm1d(C0) (affine).
// Replaces darts_tools/model_discCNN.py:350-353 / :493 behind include/gnas.h.
#include "launch.cuh"
#include "../../include/gnas.h"

namespace gnas {
namespace {
inline long long al4(long long x) { return (x + 3) / 4 * 4; }
struct StemPlan { long long acc, stat, pk, z, total; long long bacc, dmix, coef, btotal; };
StemPlan stem_plan(int B, int C0, int L) {
  StemPlan p;
  long long o = 0;
  p.acc = o; o += al4(4LL * C0);
  p.stat = o; o += al4(2LL * C0);
  p.pk = o; o += al4(4LL * 13 * C0);
  p.z = o; o += al4((long long)B * C0 * L);
  p.total = o;
  o = 0;
  p.bacc = o; o += al4(4LL * C0);
  p.dmix = o; o += 4;
  p.coef = o; o += al4(3LL * C0);
  p.btotal = o;
  return p;
}
}  // namespace
}  // namespace gnas

using namespace gnas;

extern "C" int gnas_stem_workspace(int B, int C0, int L, int64_t* fwd_floats, int64_t* bwd_floats) {
  if (B <= 0 || C0 <= 0 || C0 % 4 || L <= 0) { set_error("stem: bad shape"); return GNAS_ERR_SHAPE; }
  StemPlan p = stem_plan(B, C0, L);
  if (fwd_floats) *fwd_floats = p.total;
  if (bwd_floats) *bwd_floats = p.btotal;
  return 0;
}

extern "C" int gnas_stem_forward(int B, int C0, int L, int training, float momentum, const float* x,
                                 const float* w, const float* gamma, const float* beta, float* running,
                                 float* out, float* ws, void* stream) {
  if (B <= 0 || C0 <= 0 || C0 % 4 || L <= 0) { set_error("stem: bad shape"); return GNAS_ERR_SHAPE; }
  if (!x || !w || !gamma || !beta || !out || !ws) { set_error("stem_forward: null pointer"); return GNAS_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  StemPlan p = stem_plan(B, C0, L);
  int rc = 0;
  if (training) cudaMemsetAsync(ws + p.acc, 0, sizeof(double) * 2 * C0, st);
  PackBatch pb; pb.n = 1;
  pb.j[0].src = w; pb.j[0].dst = ws + p.pk; pb.j[0].Cout = C0; pb.j[0].Cin = 4; pb.j[0].KW = 13; pb.j[0].mode = 0;
  GNAS_LAUNCH(k_pack_weights, dim3(4, 1), dim3(256), 0, st, pb);
  DenseBatch db; db.n = 1; db.B = B; db.Cout = C0; db.pad = 6;
  DenseJob& j = db.j[0];
  j.x = x; j.xbs = 4LL * L; j.in_stat = nullptr; j.wp = ws + p.pk; j.z = ws + p.z; j.zbs = (long long)C0 * L;
  j.acc = training ? reinterpret_cast<double*>(ws + p.acc) : nullptr; j.Cin = 4; j.Lin = L; j.Lout = L; j.tin = T_NONE;
  rc |= launch_dense_fwd(13, 1, db, st);
  FinBatch f;
  f.acc = reinterpret_cast<double*>(ws + p.acc); f.stat = ws + p.stat; f.running = running;
  f.C = C0; f.training = training; f.momentum = momentum; f.nslots = 1;
  f.slot[0] = 0; f.count[0] = (float)((long long)B * L); f.rs_off[0] = 0;
  GNAS_LAUNCH(k_bn_finalize, dim3(1), dim3((C0 + 31) / 32 * 32), 0, st, f);
  BnApplyBatch ab; ab.n = 1;
  ab.j[0].z = ws + p.z; ab.j[0].y = out; ab.j[0].stat = ws + p.stat; ab.j[0].gamma = gamma; ab.j[0].beta = beta;
  ab.j[0].C = C0; ab.j[0].L = L; ab.j[0].total = (long long)B * C0 * L;
  GNAS_LAUNCH(k_bn_apply, dim3(ew_blocks(ab.j[0].total), 1), dim3(kThreads), 0, st, ab);
  if (rc) return rc;
  return check_launch("gnas_stem_forward");
}

extern "C" int gnas_stem_backward(int B, int C0, int L, const float* x, const float* w, const float* gamma,
                                  const float* dout, const float* ws, float* scratch, float* dw,
                                  float* dgamma, float* dbeta, void* stream) {
  if (B <= 0 || C0 <= 0 || C0 % 4 || L <= 0) { set_error("stem: bad shape"); return GNAS_ERR_SHAPE; }
  if (!x || !w || !gamma || !dout || !ws || !scratch || !dw || !dgamma || !dbeta) { set_error("stem_backward: null pointer"); return GNAS_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  StemPlan p = stem_plan(B, C0, L);
  (void)w;
  int rc = 0;
  const long long bs = (long long)C0 * L;
  cudaMemsetAsync(scratch, 0, sizeof(float) * (size_t)p.coef, st);
  RedBatch rb;
  rb.B = B; rb.C = C0; rb.L = L; rb.nterms = 1; rb.s1_slot = -1;
  rb.g = dout; rb.gbs = bs; rb.bacc = reinterpret_cast<double*>(scratch + p.bacc);
  rb.dmix = reinterpret_cast<double*>(scratch + p.dmix);
  rb.t[0].z = ws + p.z; rb.t[0].bs = bs; rb.t[0].mask = nullptr; rb.t[0].slot = 0; rb.t[0].widx = -1;
  GNAS_LAUNCH(k_bwd_reduce, dim3(C0, cdiv(B, 8), red_term_chunks(rb.nterms)), dim3(kThreads), 0, st, rb);
  CoefBatch cb;
  cb.bacc = reinterpret_cast<const double*>(scratch + p.bacc); cb.stat = ws + p.stat; cb.coef = scratch + p.coef;
  cb.mixw = nullptr; cb.dmix = reinterpret_cast<double*>(scratch + p.dmix);
  cb.gamma = gamma; cb.dgamma = dgamma; cb.dbeta = dbeta;
  cb.C = C0; cb.nslots = 1; cb.s1_slot = -1; cb.slot[0] = 0; cb.widx[0] = -1; cb.count[0] = (float)((long long)B * L);
  GNAS_LAUNCH(k_bn_bwd_coef, dim3(1), dim3((C0 + 31) / 32 * 32), 0, st, cb);
  WgBatch gb; gb.n = 1; gb.B = B; gb.Cout = C0;
  WgJob& wj = gb.j[0];
  wj.g = dout; wj.gbs = bs; wj.z = ws + p.z; wj.zbs = bs; wj.coef = scratch + p.coef;
  wj.r = x; wj.rbs = 4LL * L; wj.r_stat = nullptr; wj.dw = dw; wj.Cin = 4; wj.Lo = L; wj.Lin = L; wj.tin = T_NONE;
  rc |= launch_conv_wgrad(13, 1, 6, gb, st);
  if (rc) return rc;
  return check_launch("gnas_stem_backward");
}

// ---- on-device data path: one byte per base instead of the fp32 one-hot [B][4][L] (16x fewer bytes over PCIe).
// codes[b][l] in {0, 1, 2, 3} = the hot channel, anything else (e.g. 255 for N) = all-zero column.
// Restates what the reference's loader materialises on the host (data_preprocessing_TF.py:90-149).
namespace gnas {
static __global__ void __launch_bounds__(256) k_onehot_expand(const unsigned char* __restrict__ codes, float* __restrict__ x,
                                                              long long B, long long L) {
  const long long total = B * L;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / L, l = i - b * L;
    const int c = codes[i];
    float* xb = x + b * 4 * L + l;
#pragma unroll
    for (int k = 0; k < 4; ++k) xb[k * L] = c == k ? 1.f : 0.f;
  }
}
}  // namespace gnas

extern "C" int gnas_onehot_expand(const uint8_t* codes, int64_t B, int64_t L, float* x, void* stream) {
  if (!codes || !x || B < 1 || L < 1) { set_error("onehot_expand: bad argument"); return GNAS_ERR_ARG; }
  GNAS_LAUNCH(k_onehot_expand, dim3(ew_blocks(B * L)), dim3(256), 0, (cudaStream_t)stream, codes, x, (long long)B, (long long)L);
  return check_launch("gnas_onehot_expand");
}

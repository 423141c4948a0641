// CNN search cell / MixedOp: workspace plan, phase orchestration, C ABI.
// Replaces darts_tools/model_discCNN.py:46-201 (MixedOp, CNN_Cell_search) and the
// operator zoo generalNAS_tools/operations_14_9.py behind include/gnas.h.
#include "launch.cuh"
#include "conv15_tc.cuh"
#include "wgrad_tc.cuh"
#include <cstdlib>
#include "../../include/gnas.h"

namespace gnas {

static thread_local char g_err[512] = "";
void set_error(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }
const char* get_error() { return g_err; }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

namespace {

constexpr int kNodeE0[4] = {0, 2, 5, 9};

struct OpPlan {
  bool on = false;
  int q = -1;               // column of the mixing-weight row
  long long t[4] = {-1, -1, -1, -1};   // forward tensors in ws
  int slot[2] = {-1, -1};   // BN slots
  long long pk[2] = {-1, -1};          // packed forward weights in ws
  long long pkb[2] = {-1, -1};         // packed backward weights in scratch (conv_15; pointwise on the tensor-core path)
};
struct EdgePlan {
  int node = 0, src = 0, stride = 1, Lin = 0;
  OpPlan op[GNAS_N_PRIMS];
};
struct CellPlan {
  int n_edges = 0, n_nodes = 0, nslots = 0;
  long long nCL_out = 0, nCL_p = 0;   // B*C*L_out, B*C*L_p
  EdgePlan e[GNAS_MAX_EDGES];
  int node_slot[4] = {-1, -1, -1, -1};
  // forward workspace (floats)
  long long acc = 0, stat = 0, pre_z[2] = {0, 0}, pre_y[2] = {0, 0}, pre_pk[2] = {0, 0};
  long long fwd_total = 0;
  // backward scratch (floats)
  long long bacc = 0, dmix = 0, coef = 0, G[4] = {0, 0, 0, 0}, dS[2] = {0, 0};
  long long tmp[5][9];      // per edge-in-phase temporaries
  long long bwd_total = 0;
};

inline long long al4(long long x) { return (x + 3) / 4 * 4; }

// tcgen05 path for the stride-1 conv_15 / pointwise convolutions (conv15_tc.cuh); GNAS_TC=0 selects the SIMT kernels
inline bool tc_on(int C, int stride) {
#ifdef GNAS_EMUL
  (void)C; (void)stride; return false;
#else
  return tc_enabled() && stride == 1 && tc::tc_supported(C);
#endif
}
// debugging aid: GNAS_TC_PART bit 0 = conv_15 forward, bit 1 / 2 = data gradient of its second / first conv,
// bit 3 = pointwise convs (forward and data gradient), bit 4 = pointwise weight gradient,
// bit 5 = conv_15 weight gradient, bit 6 = strided conv_15 data gradient, bit 7 = strided conv_15 forward,
// bit 8 = strided conv_15 weight gradient
inline bool tc_part(int bit) {
  static int v = -1;
  if (v < 0) { const char* s = getenv("GNAS_TC_PART"); v = s ? atoi(s) : 511; }
  return (v >> bit) & 1;
}
#ifndef GNAS_EMUL
// the same jobs, handed to the tensor-core kernel
int launch_conv_fwd_tc(DenseBatch& db, int KW, cudaStream_t st) {
  tc::TcBatch tb; tb.n = 0; tb.B = db.B; tb.C = db.Cout; tb.KW = KW;
  for (int i = 0; i < db.n; ++i) {
    const DenseJob& j = db.j[i];
    tc::TcJob& t = tb.j[tb.n++];
    t.x = j.x; t.xbs = j.xbs; t.x2 = nullptr; t.x2bs = 0; t.in_stat = j.in_stat; t.w = j.wp;
    t.z = j.z; t.zbs = j.zbs; t.acc = j.acc; t.mref = nullptr; t.mrbs = 0; t.mstat = nullptr;
    t.L = j.Lout; t.tin = j.tin; t.omode = O_STORE;
    t.zl = t.L; t.zs = 1; t.zo = 0; t.halo = -1; t.xl = t.L; t.xs = 1; t.xo = 0;
  }
  db.n = 0;
  return tc::launch_conv_tc(tb, st);
}
int launch_conv_dx_tc(DxBatch& xb, int KW, cudaStream_t st) {
  tc::TcBatch tb; tb.n = 0; tb.B = xb.B; tb.C = xb.Cin; tb.KW = KW;
  for (int i = 0; i < xb.n; ++i) {
    const DxJob& j = xb.j[i];
    tc::TcJob& t = tb.j[tb.n++];
    t.x = j.g; t.xbs = j.gbs; t.x2 = j.z; t.x2bs = j.zbs; t.in_stat = j.coef; t.w = j.wt;
    t.z = j.out; t.zbs = j.obs; t.acc = j.omode == O_BNRELU_STORE ? j.bacc : nullptr;
    t.mref = j.xin; t.mrbs = j.xibs; t.mstat = j.xin_stat;
    t.L = j.Lo; t.tin = -1; t.omode = j.omode;
    t.zl = t.L; t.zs = 1; t.zo = 0; t.halo = -1; t.xl = t.L; t.xs = 1; t.xo = 0;
  }
  xb.n = 0;
  return tc::launch_conv_tc(tb, st);
}
// Data gradient of a stride-S conv_15: the outputs m = S q + rho of each residue class rho are a stride-1
// correlation over dz with the taps t = (rho + 7) mod S + S k (5 per class for S = 3; 7 / 8, padded to 8, for S = 2);
// one tensor-core job per (conv, class), weights packed per class by k_pack_tc mode 2.
// Forward of a stride-S conv_15: the input is split into its S phases x[S q + phi]; each phase is a stride-1
// correlation with the taps t = (phi + 7) mod S + S k, and the phases are accumulated by S consecutive launches
// (store, then add; the last one also takes the BN statistics of the finished sum).
int launch_conv_fwd_strided_tc(DenseBatch& db, int S, cudaStream_t st);
inline int strided_dx_taps(int S) { return S == 3 ? 5 : 8; }
inline long long strided_dx_pack_floats(int C, int S) { return 2LL * C * C * strided_dx_taps(S) * S; }
int launch_conv_fwd_strided_tc(DenseBatch& db, int S, cudaStream_t st) {
  const int NT = strided_dx_taps(S);
  int rc = 0;
  for (int phi = 0; phi < S; ++phi) {
    tc::TcBatch tb; tb.n = 0; tb.B = db.B; tb.C = db.Cout; tb.KW = NT;
    const int t0 = (phi + 7) % S;
    for (int i = 0; i < db.n; ++i) {
      const DenseJob& j = db.j[i];
      tc::TcJob& t = tb.j[tb.n++];
      t.x = j.x; t.xbs = j.xbs; t.x2 = nullptr; t.x2bs = 0; t.in_stat = j.in_stat;
      t.w = j.wp + (long long)phi * 2 * db.Cout * db.Cout * NT;
      t.z = j.z; t.zbs = j.zbs; t.acc = phi == S - 1 ? j.acc : nullptr; t.mref = nullptr; t.mrbs = 0; t.mstat = nullptr;
      t.L = j.Lout; t.tin = j.tin; t.omode = phi == 0 ? O_STORE : O_ACC;
      t.zl = t.L; t.zs = 1; t.zo = 0; t.halo = (7 + phi - t0) / S; t.xl = j.Lin; t.xs = S; t.xo = phi;
    }
    rc |= tc::launch_conv_tc(tb, st);
  }
  db.n = 0;
  return rc;
}
int launch_conv_dx_strided_tc(DxBatch& xb, int S, cudaStream_t st) {
  const int NT = strided_dx_taps(S);
  tc::TcBatch tb; tb.n = 0; tb.B = xb.B; tb.C = xb.Cin; tb.KW = NT;
  int rc = 0;
  for (int i = 0; i < xb.n; ++i) {
    const DxJob& j = xb.j[i];
    for (int rho = 0; rho < S; ++rho) {
      const int t0 = (rho + 7) % S, D0 = (rho + 7 - t0) / S;
      tc::TcJob& t = tb.j[tb.n++];
      t.x = j.g; t.xbs = j.gbs; t.x2 = j.z; t.x2bs = j.zbs; t.in_stat = j.coef;
      t.w = j.wt + (long long)rho * 2 * xb.Cin * xb.Cin * NT;
      t.z = j.out; t.zbs = j.obs; t.acc = nullptr;
      t.mref = j.xin; t.mrbs = j.xibs; t.mstat = nullptr;
      t.L = j.Lo; t.tin = -1; t.omode = j.omode;
      t.zl = j.Lin; t.zs = S; t.zo = rho; t.halo = NT - 1 - D0; t.xl = t.L; t.xs = 1; t.xo = 0;
      if (tb.n == kMaxDense) rc |= tc::launch_conv_tc(tb, st);
    }
  }
  xb.n = 0;
  rc |= tc::launch_conv_tc(tb, st);
  return rc;
}
void pack_tc(tc::TcPackBatch& pb, cudaStream_t st) {
  if (pb.n == 0) return;
  GNAS_LAUNCH(tc::k_pack_tc, dim3(32, pb.n), dim3(256), 0, st, pb);
  pb.n = 0;
}
#else
int launch_conv_fwd_tc(DenseBatch&, int, cudaStream_t) { return GNAS_ERR_UNSUPPORTED; }
int launch_conv_dx_tc(DxBatch&, int, cudaStream_t) { return GNAS_ERR_UNSUPPORTED; }
int launch_conv_dx_strided_tc(DxBatch&, int, cudaStream_t) { return GNAS_ERR_UNSUPPORTED; }
int launch_conv_fwd_strided_tc(DenseBatch&, int, cudaStream_t) { return GNAS_ERR_UNSUPPORTED; }
inline long long strided_dx_pack_floats(int, int) { return 0; }
#endif
inline bool tc_fwd_on(int C, int stride) { return tc_on(C, stride) && tc_part(0); }
inline bool tc_dx_on(int C, int stride, int conv) { return tc_on(C, stride) && tc_part(conv == 1 ? 1 : 2); }
inline bool tc_sdx_on(int C, int S) {       // strided conv_15 data gradient: instantiated for the supernet's reduction cells
#ifdef GNAS_EMUL
  (void)C; (void)S; return false;
#else
  return tc_enabled() && tc_part(6) && ((S == 2 && (C == 32 || C == 64)) || (S == 3 && C == 128));
#endif
}
inline bool tc_sfw_on(int C, int S) { return tc_sdx_on(C, S) && tc_part(7); }
inline bool tc_pw_on(int C) { return tc_on(C, 1) && tc_part(3); }
inline long long pw_pack_floats(int C) { return (long long)C * C * (tc_pw_on(C) ? 2 : 1); }
inline long long conv15_pack_floats(int C) { return (long long)C * C * 15 * (tc_on(C, 1) ? 2 : 1); }

int validate(const GnasCellDesc* d) {
  if (!d) { set_error("null descriptor"); return GNAS_ERR_ARG; }
  if (d->B <= 0 || d->C <= 0 || d->C % 4 != 0) { set_error("C must be a positive multiple of 4"); return GNAS_ERR_SHAPE; }
  if (d->stride < 1 || d->stride > 3) { set_error("stride must be 1, 2 or 3"); return GNAS_ERR_SHAPE; }
  if (d->has_pre ? d->n_edges != 14 : d->n_edges != 1) { set_error("n_edges must be 14 (cell) or 1 (MixedOp)"); return GNAS_ERR_SHAPE; }
  if (d->L_out != (d->L_p - 1) / d->stride + 1) { set_error("L_out != (L_p-1)/stride+1"); return GNAS_ERR_SHAPE; }
  if (d->has_pre) {
    if (d->C_pp % 4 || d->C_p % 4) { set_error("input channels must be multiples of 4"); return GNAS_ERR_SHAPE; }
    const int l0 = d->pre0_factorized ? (d->L_pp - 1) / 2 + 1 : d->L_pp;
    if (l0 != d->L_p) { set_error("preprocess0 output length != L_p"); return GNAS_ERR_SHAPE; }
  }
  for (int e = 0; e < d->n_edges; ++e) {
    int n = 0;
    for (int k = 0; k < GNAS_N_PRIMS; ++k) n += d->sw[e][k] ? 1 : 0;
    // DARTS / P-DARTS / CWP-DARTS enable exactly n_prims per edge; the mask supernets (OSP-NAS, Hyperband, random
    // search) enable any subset, an edge may even be empty (it then contributes zero)
    if (n > d->n_prims) { set_error("an edge enables more primitives than the mixing-weight rows have columns"); return GNAS_ERR_SHAPE; }
  }
  return 0;
}

void make_plan(const GnasCellDesc* d, CellPlan& P) {
  const int C = d->C;
  P.n_edges = d->n_edges;
  P.n_nodes = d->has_pre ? 4 : 1;
  P.nCL_out = (long long)d->B * C * d->L_out;
  P.nCL_p = (long long)d->B * C * d->L_p;
  int slot = d->has_pre ? 2 : 0;
  // edges
  for (int e = 0; e < d->n_edges; ++e) {
    EdgePlan& E = P.e[e];
    if (d->has_pre) {
      int i = 3;
      while (kNodeE0[i] > e) --i;
      E.node = i; E.src = e - kNodeE0[i];
    } else { E.node = 0; E.src = 1; }
    E.stride = (E.src < 2) ? d->stride : 1;
    E.Lin = (E.src < 2) ? d->L_p : d->L_out;
    int q = 0;
    for (int k = 0; k < GNAS_N_PRIMS; ++k) {
      OpPlan& O = E.op[k];
      O.on = d->sw[e][k] != 0;
      if (!O.on) continue;
      O.q = q++;
      switch (k) {
        case GNAS_MAX_POOL: case GNAS_AVG_POOL: O.slot[0] = slot++; break;
        case GNAS_SKIP: if (E.stride > 1) O.slot[0] = slot++; break;
        case GNAS_SEP15: case GNAS_SEP9: case GNAS_CONV15: O.slot[0] = slot++; O.slot[1] = slot++; break;
        case GNAS_DIL15: case GNAS_DIL9: O.slot[0] = slot++; break;
        default: break;
      }
    }
  }
  for (int i = 0; i < P.n_nodes; ++i) P.node_slot[i] = slot++;
  P.nslots = slot;
  // ---- forward workspace
  long long o = 0;
  P.acc = o; o += al4((long long)P.nslots * 4 * C);       // doubles: nslots*2*C
  P.stat = o; o += al4((long long)P.nslots * 2 * C);
  if (d->has_pre) {
    for (int i = 0; i < 2; ++i) { P.pre_z[i] = o; o += al4(P.nCL_p); }
    for (int i = 0; i < 2; ++i) { P.pre_y[i] = o; o += al4(P.nCL_p); }
    P.pre_pk[0] = o; o += al4((long long)C * d->C_pp);
    P.pre_pk[1] = o; o += al4((long long)C * d->C_p);
  }
  for (int e = 0; e < d->n_edges; ++e) {
    EdgePlan& E = P.e[e];
    for (int k = 0; k < GNAS_N_PRIMS; ++k) {
      OpPlan& O = E.op[k];
      if (!O.on) continue;
      int nt = 0, npk[2] = {0, 0};
      switch (k) {
        case GNAS_MAX_POOL: case GNAS_AVG_POOL: nt = 1; break;
        case GNAS_SKIP: if (E.stride > 1) { nt = 1; npk[0] = C * C; } break;
        case GNAS_SEP15: case GNAS_SEP9: nt = 4; npk[0] = npk[1] = (int)pw_pack_floats(C); break;
        case GNAS_DIL15: case GNAS_DIL9: nt = 2; npk[0] = (int)pw_pack_floats(C); break;
        case GNAS_CONV15:
          nt = 2; npk[0] = npk[1] = (int)conv15_pack_floats(C);
          if (tc_sfw_on(C, E.stride) && strided_dx_pack_floats(C, E.stride) > npk[0]) npk[0] = (int)strided_dx_pack_floats(C, E.stride);
          break;
        default: break;
      }
      for (int i = 0; i < nt; ++i) { O.t[i] = o; o += al4(P.nCL_out); }
      for (int i = 0; i < 2; ++i) if (npk[i]) { O.pk[i] = o; o += al4(npk[i]); }
    }
  }
  P.fwd_total = o;
  // ---- backward scratch
  o = 0;
  P.bacc = o; o += al4((long long)P.nslots * 4 * C);
  P.dmix = o; o += al4(2LL * GNAS_MAX_EDGES * GNAS_N_PRIMS);
  P.coef = o; o += al4((long long)P.nslots * 3 * C);
  if (d->has_pre) {
    for (int i = 0; i < 4; ++i) { P.G[i] = o; o += al4(P.nCL_out); }
    for (int i = 0; i < 2; ++i) { P.dS[i] = o; o += al4(P.nCL_p); }
  }
  const int max_phase_edges = d->has_pre ? 5 : 1;
  for (int s = 0; s < max_phase_edges; ++s)
    for (int i = 0; i < 9; ++i) { P.tmp[s][i] = o; o += al4(P.nCL_out); }
  for (int e = 0; e < d->n_edges; ++e) {
    OpPlan& O = P.e[e].op[GNAS_CONV15];
    if (O.on) for (int i = 0; i < 2; ++i) {
      long long n = conv15_pack_floats(C);
      if (i == 0 && tc_sdx_on(C, P.e[e].stride)) { const long long m = strided_dx_pack_floats(C, P.e[e].stride); n = m > n ? m : n; }
      O.pkb[i] = o; o += al4(n);
    }
    if (tc_pw_on(C)) {   // tensor-core data gradient of the pointwise convs: transposed hi | lo packs
      const int ks[4] = {GNAS_SEP15, GNAS_SEP9, GNAS_DIL15, GNAS_DIL9};
      for (int v = 0; v < 4; ++v) {
        OpPlan& Q = P.e[e].op[ks[v]];
        if (!Q.on) continue;
        for (int i = 0; i < (v < 2 ? 2 : 1); ++i) { Q.pkb[i] = o; o += al4(2LL * C * C); }
      }
    }
  }
  P.bwd_total = o;
}

struct Ctx {
  const GnasCellDesc* d;
  CellPlan P;
  const float* params; float* grads; float* running;
  const float* s0; const float* s1; const float* mixw;
  const int64_t* skip_mask;
  float* out; float* ws; float* scratch;
  cudaStream_t st;
  // Independent chains of a cell pass run beside the main chain on side streams (fork / join by events; a CUDA-graph
  // capture records them as plain dependencies): 1 = weight-gradient kernels (leaves of the backward pass) and the
  // conv_15 forward chain, 2 = pools / strided skip (forward) and the pool / skip data gradients (backward),
  // 3 = the data gradient of the second conv_15 convolution.
  SideStream side[4] = {{nullptr, nullptr, {nullptr, nullptr}}, {nullptr, nullptr, {nullptr, nullptr}},
                        {nullptr, nullptr, {nullptr, nullptr}}, {nullptr, nullptr, {nullptr, nullptr}}};
  bool forked[4] = {false, false, false, false};
  int rc = 0;
  void init_sides(bool on) { for (int w = 1; w < 4; ++w) side[w] = side_stream(on ? w : -1); }
  // stream for a launch whose inputs are complete on `st` at this point
  cudaStream_t fork(int w) {
    if (!side[w].st) return st;
    cudaEventRecord(side[w].fork, st);
    cudaStreamWaitEvent(side[w].st, side[w].fork, 0);
    forked[w] = true;
    return side[w].st;
  }
  // before anything the forked launches read is overwritten / anything they write is consumed, and before the call returns
  void join(int w) {
    if (!forked[w]) return;
    cudaEventRecord(side[w].join[0], side[w].st);
    cudaStreamWaitEvent(st, side[w].join[0], 0);
    forked[w] = false;
  }
  cudaStream_t wg_stream() { return fork(1); }
  void wg_join() { join(1); }

  const float* W(int e, int k, int i) const { return params + d->w_off[e][k][i]; }
  float* dW(int e, int k, int i) const { return grads + d->w_off[e][k][i]; }
  float* T(int e, int k, int i) const { return ws + P.e[e].op[k].t[i]; }
  double* acc(int slot) const { return reinterpret_cast<double*>(ws + P.acc) + (long long)slot * 2 * d->C; }
  float* stat(int slot) const { return ws + P.stat + (long long)slot * 2 * d->C; }
  double* bacc(int slot) const { return reinterpret_cast<double*>(scratch + P.bacc) + (long long)slot * 2 * d->C; }
  float* coef(int slot) const { return scratch + P.coef + (long long)slot * 3 * d->C; }
  double* dmix() const { return reinterpret_cast<double*>(scratch + P.dmix); }
  int widx(int e, int k) const { return e * d->n_prims + P.e[e].op[k].q; }
  const float* mask(int e) const { return skip_mask ? reinterpret_cast<const float*>((uintptr_t)skip_mask[e]) : nullptr; }
  // source state of an edge: pointer + batch stride
  void src(int e, const float*& p, long long& bs) const {
    const EdgePlan& E = P.e[e];
    if (!d->has_pre) { p = s1; bs = (long long)d->C * d->L_p; return; }
    if (E.src < 2) { p = ws + P.pre_y[E.src]; bs = (long long)d->C * d->L_p; }
    else { p = out + (long long)(E.src - 2) * d->C * d->L_out; bs = 4LL * d->C * d->L_out; }
  }
  void node_out(int node, float*& p, long long& bs) const {
    if (!d->has_pre) { p = out; bs = (long long)d->C * d->L_out; }
    else { p = out + (long long)node * d->C * d->L_out; bs = 4LL * d->C * d->L_out; }
  }
};

void finalize(Ctx& c, const int* slots, const float* counts, const long long* rs, int n) {
  if (n == 0) return;
  FinBatch f;
  f.acc = reinterpret_cast<double*>(c.ws + c.P.acc);
  f.stat = c.ws + c.P.stat;
  f.running = c.running;
  f.C = c.d->C; f.training = c.d->training; f.momentum = c.d->bn_momentum;
  f.nslots = n;
  for (int i = 0; i < n; ++i) { f.slot[i] = (short)slots[i]; f.count[i] = counts[i]; f.rs_off[i] = rs[i]; }
  GNAS_LAUNCH(k_bn_finalize, dim3(n), dim3(c.d->C < 256 ? (c.d->C + 31) / 32 * 32 : 256), 0, c.st, f);
}

// -------------------------------------------------------------------------- forward of a set of edges
void edges_forward(Ctx& c, const int* edges, int ne) {
  const GnasCellDesc* d = c.d;
  const int C = d->C, B = d->B;
  const int S = c.P.e[edges[0]].stride;
  const int Lin = c.P.e[edges[0]].Lin, Lo = d->L_out;
  const long long obs = (long long)C * Lo;
  const bool train = d->training != 0;
  int fs[kMaxSlots]; float fc[kMaxSlots]; long long fr[kMaxSlots]; int nf = 0;
  auto fin_add = [&](int e, int k, int which) {
    fs[nf] = c.P.e[e].op[k].slot[which]; fc[nf] = (float)((long long)B * Lo); fr[nf] = d->bn_off[e][k][which]; ++nf;
  };
  // the conv_15 chain reads only the edge sources: it runs on the side stream, forked here, beside pools / depthwise / pointwise
  cudaStream_t cst = c.fork(1);
  cudaStream_t pst = c.fork(2);       // pools and the strided skip: short kernels on the edge sources, beside both chains
  // ---- stage 1
  {
    PoolBatch pb; pb.n = 0; pb.B = B; pb.C = C; pb.S = S;
    for (int i = 0; i < ne; ++i) {
      const int e = edges[i];
      const OpPlan& mx = c.P.e[e].op[GNAS_MAX_POOL]; const OpPlan& av = c.P.e[e].op[GNAS_AVG_POOL];
      if (!mx.on && !av.on) continue;
      PoolJob& j = pb.j[pb.n++];
      c.src(e, j.x, j.xbs);
      j.zmax = mx.on ? c.T(e, GNAS_MAX_POOL, 0) : nullptr;
      j.zavg = av.on ? c.T(e, GNAS_AVG_POOL, 0) : nullptr;
      j.acc_max = (mx.on && train) ? c.acc(mx.slot[0]) : nullptr;
      j.acc_avg = (av.on && train) ? c.acc(av.slot[0]) : nullptr;
      j.Lin = Lin; j.Lout = Lo;
      if (mx.on) fin_add(e, GNAS_MAX_POOL, 0);
      if (av.on) fin_add(e, GNAS_AVG_POOL, 0);
      if (pb.n == kMaxPool) launch_pool_fwd(pb, pst);
    }
    launch_pool_fwd(pb, pst);
  }
  DenseBatch db; db.n = 0; db.B = B; db.Cout = C; db.pad = 0;
  if (S > 1) {  // strided skip = FactorizedReduce (both halves read the same positions: quirk Q4)
    for (int i = 0; i < ne; ++i) {
      const int e = edges[i];
      const OpPlan& O = c.P.e[e].op[GNAS_SKIP];
      if (!O.on) continue;
      DenseJob& j = db.j[db.n++];
      c.src(e, j.x, j.xbs); j.in_stat = nullptr; j.wp = c.ws + O.pk[0];
      j.z = c.T(e, GNAS_SKIP, 0); j.zbs = obs; j.acc = train ? c.acc(O.slot[0]) : nullptr;
      j.Cin = C; j.Lin = Lin; j.Lout = Lo; j.tin = T_RELU;
      fin_add(e, GNAS_SKIP, 0);
    }
    c.rc |= launch_dense_fwd(1, S, db, pst);
  }
  auto pw_fwd = [&](DenseBatch& b) { return tc_pw_on(C) ? launch_conv_fwd_tc(b, 1, c.st) : launch_dense_fwd(1, 1, b, c.st); };
  // depthwise stage 1 (sep: dilation 1, dil: dilation 2)
  const int dwk[4] = {GNAS_SEP15, GNAS_SEP9, GNAS_DIL15, GNAS_DIL9};
  const int dwK[4] = {15, 9, 15, 9}, dwD[4] = {1, 1, 2, 2}, dwP[4] = {7, 4, 14, 8};
  {   // fused: one pass over relu(x) per edge for all four candidates
    EdgeFwdBatch eb; eb.n = 0; eb.B = B; eb.C = C;
    for (int i = 0; i < ne; ++i) {
      const int e = edges[i];
      bool any = false;
      for (int v = 0; v < 4; ++v) any = any || c.P.e[e].op[dwk[v]].on;
      if (!any) continue;
      EdgeFwdJob& j = eb.j[eb.n++];
      c.src(e, j.x, j.xbs); j.ubs = obs; j.Lin = Lin; j.Lout = Lo;
      for (int v = 0; v < 4; ++v) {
        const bool on = c.P.e[e].op[dwk[v]].on;
        j.wd[v] = on ? c.W(e, dwk[v], 0) : nullptr;
        j.u[v] = on ? c.T(e, dwk[v], 0) : nullptr;
      }
      if (eb.n == kMaxEdgeFwd) c.rc |= launch_edge_fwd(S, eb, c.st);
    }
    c.rc |= launch_edge_fwd(S, eb, c.st);
  }
  // pointwise stage 1
  for (int i = 0; i < ne; ++i) {
    const int e = edges[i];
    for (int v = 0; v < 4; ++v) {
      const OpPlan& O = c.P.e[e].op[dwk[v]];
      if (!O.on) continue;
      DenseJob& j = db.j[db.n++];
      j.x = c.T(e, dwk[v], 0); j.xbs = obs; j.in_stat = nullptr; j.wp = c.ws + O.pk[0];
      j.z = c.T(e, dwk[v], 1); j.zbs = obs; j.acc = train ? c.acc(O.slot[0]) : nullptr;
      j.Cin = C; j.Lin = Lo; j.Lout = Lo; j.tin = T_NONE;
      fin_add(e, dwk[v], 0);
      if (db.n == kMaxDense) c.rc |= pw_fwd(db);
    }
  }
  c.rc |= pw_fwd(db);
  // conv_15 stage 1 (independent of the depthwise / pointwise chain: runs beside it on the side stream; the stream
  // was forked at the top of the function, the kernels above were enqueued on c.st in the meantime)
  db.pad = 7;
  for (int i = 0; i < ne; ++i) {
    const int e = edges[i];
    const OpPlan& O = c.P.e[e].op[GNAS_CONV15];
    if (!O.on) continue;
    DenseJob& j = db.j[db.n++];
    c.src(e, j.x, j.xbs); j.in_stat = nullptr; j.wp = c.ws + O.pk[0];
    j.z = c.T(e, GNAS_CONV15, 0); j.zbs = obs; j.acc = train ? c.acc(O.slot[0]) : nullptr;
    j.Cin = C; j.Lin = Lin; j.Lout = Lo; j.tin = T_RELU;
    fin_add(e, GNAS_CONV15, 0);
  }
  c.rc |= tc_fwd_on(C, S) ? launch_conv_fwd_tc(db, 15, cst)
          : (S > 1 && tc_sfw_on(C, S)) ? launch_conv_fwd_strided_tc(db, S, cst) : launch_dense_fwd(15, S, db, cst);
  c.join(1);
  c.join(2);
  finalize(c, fs, fc, fr, nf);
  nf = 0;
  cst = c.fork(1);
  // ---- stage 2 (sep_conv second half, conv_15 second conv) reads BN+ReLU of stage 1 on the fly
  {
    DwBatch wb; wb.n = 0; wb.B = B; wb.C = C; wb.pad = 0;
    for (int v = 0; v < 2; ++v)
      for (int i = 0; i < ne; ++i) {
        const int e = edges[i];
        const OpPlan& O = c.P.e[e].op[dwk[v]];
        if (!O.on) continue;
        DwJob& j = wb.j[wb.n++];
        j.x = c.T(e, dwk[v], 1); j.xbs = obs; j.in_stat = c.stat(O.slot[0]); j.wd = c.W(e, dwk[v], 2);
        j.u = c.T(e, dwk[v], 2); j.ubs = obs; j.Lin = Lo; j.Lout = Lo; j.tin = T_BN_RELU; j.K = dwK[v];
        if (wb.n == kMaxDw) c.rc |= launch_dw2_fwd(wb, c.st);
      }
    c.rc |= launch_dw2_fwd(wb, c.st);
  }
  db.pad = 0;
  for (int i = 0; i < ne; ++i) {
    const int e = edges[i];
    for (int v = 0; v < 2; ++v) {
      const OpPlan& O = c.P.e[e].op[dwk[v]];
      if (!O.on) continue;
      DenseJob& j = db.j[db.n++];
      j.x = c.T(e, dwk[v], 2); j.xbs = obs; j.in_stat = nullptr; j.wp = c.ws + O.pk[1];
      j.z = c.T(e, dwk[v], 3); j.zbs = obs; j.acc = train ? c.acc(O.slot[1]) : nullptr;
      j.Cin = C; j.Lin = Lo; j.Lout = Lo; j.tin = T_NONE;
      fin_add(e, dwk[v], 1);
    }
  }
  c.rc |= pw_fwd(db);
  db.pad = 7;
  for (int i = 0; i < ne; ++i) {
    const int e = edges[i];
    const OpPlan& O = c.P.e[e].op[GNAS_CONV15];
    if (!O.on) continue;
    DenseJob& j = db.j[db.n++];
    j.x = c.T(e, GNAS_CONV15, 0); j.xbs = obs; j.in_stat = c.stat(O.slot[0]); j.wp = c.ws + O.pk[1];
    j.z = c.T(e, GNAS_CONV15, 1); j.zbs = obs; j.acc = train ? c.acc(O.slot[1]) : nullptr;
    j.Cin = C; j.Lin = Lo; j.Lout = Lo; j.tin = T_BN_RELU;
    fin_add(e, GNAS_CONV15, 1);
  }
  c.rc |= tc_fwd_on(C, 1) ? launch_conv_fwd_tc(db, 15, cst) : launch_dense_fwd(15, 1, db, cst);
  c.join(1);
  finalize(c, fs, fc, fr, nf);
}

// final (mixed) tensor + slot of a primitive, or -1 when it has no BN output
inline int final_tensor(int k, int stride) {
  switch (k) {
    case GNAS_MAX_POOL: case GNAS_AVG_POOL: return 0;
    case GNAS_SKIP: return stride > 1 ? 0 : -1;
    case GNAS_SEP15: case GNAS_SEP9: return 3;
    case GNAS_DIL15: case GNAS_DIL9: return 1;
    case GNAS_CONV15: return 1;
    default: return -1;
  }
}
inline int final_slot_idx(int k) { return (k == GNAS_SEP15 || k == GNAS_SEP9 || k == GNAS_CONV15) ? 1 : 0; }

void node_mix(Ctx& c, int node) {
  const GnasCellDesc* d = c.d;
  MixBatch m;
  m.B = d->B; m.C = d->C; m.L = d->L_out; m.nterms = 0; m.mixw = c.mixw;
  c.node_out(node, m.out, m.obs);
  for (int e = 0; e < c.P.n_edges; ++e) {
    if (c.P.e[e].node != node) continue;
    for (int k = 1; k < GNAS_N_PRIMS; ++k) {
      const OpPlan& O = c.P.e[e].op[k];
      if (!O.on) continue;
      MixTerm& t = m.t[m.nterms++];
      const int ft = final_tensor(k, c.P.e[e].stride);
      t.widx = c.widx(e, k); t.mask = nullptr;
      if (ft >= 0) { t.z = c.T(e, k, ft); t.bs = (long long)d->C * d->L_out; t.stat = c.stat(O.slot[final_slot_idx(k)]); }
      else { c.src(e, t.z, t.bs); t.stat = nullptr; t.mask = c.mask(e); }
    }
  }
  GNAS_LAUNCH(k_mix_fwd, dim3(ew_blocks((long long)m.B * m.C * m.L)), dim3(kThreads), 0, c.st, m);
}

void pack_forward(Ctx& c) {
  const GnasCellDesc* d = c.d;
  const int C = d->C;
  PackBatch pb; pb.n = 0;
#ifndef GNAS_EMUL
  tc::TcPackBatch tp; tp.n = 0;
#endif
  auto add = [&](const float* src, float* dst, int co, int ci, int kw, int mode) {
    PackJob& j = pb.j[pb.n++];
    j.src = src; j.dst = dst; j.Cout = co; j.Cin = ci; j.KW = kw; j.mode = mode;
    if (pb.n == kMaxPack) { GNAS_LAUNCH(k_pack_weights, dim3(8, pb.n), dim3(256), 0, c.st, pb); pb.n = 0; }
  };
  if (d->has_pre) {
    add(c.params + d->pre_w_off[0], c.ws + c.P.pre_pk[0], C, d->C_pp, 1, 0);
    add(c.params + d->pre_w_off[1], c.ws + c.P.pre_pk[1], C, d->C_p, 1, 0);
  }
  for (int e = 0; e < c.P.n_edges; ++e) {
    const EdgePlan& E = c.P.e[e];
    for (int k = 0; k < GNAS_N_PRIMS; ++k) {
      const OpPlan& O = E.op[k];
      if (!O.on) continue;
      if (k == GNAS_SKIP && E.stride > 1) add(c.W(e, k, 0), c.ws + O.pk[0], C, C, 1, 0);
      if (k == GNAS_SEP15 || k == GNAS_SEP9 || k == GNAS_DIL15 || k == GNAS_DIL9) {
        const int npw = (k == GNAS_SEP15 || k == GNAS_SEP9) ? 2 : 1;
        for (int i = 0; i < npw; ++i) {
          const float* src = c.W(e, k, i == 0 ? 1 : 3);
          if (!tc_pw_on(C)) { add(src, c.ws + O.pk[i], C, C, 1, 0); continue; }
#ifndef GNAS_EMUL
          tc::TcPackJob& t = tp.j[tp.n++];
          t.src = src; t.dst = c.ws + O.pk[i]; t.C = C; t.mode = 0; t.KW = 1;
          if (tp.n == 64) pack_tc(tp, c.st);
#endif
        }
      }
      if (k == GNAS_CONV15) {
        for (int i = 0; i < 2; ++i) {
          const int cs = i == 0 ? E.stride : 1;
#ifndef GNAS_EMUL
          if (i == 0 && cs > 1 && tc_sfw_on(C, cs)) {
            const int NT = strided_dx_taps(cs);
            for (int phi = 0; phi < cs; ++phi) {
              tc::TcPackJob& t = tp.j[tp.n++];
              t.src = c.W(e, k, 0); t.dst = c.ws + O.pk[0] + (long long)phi * 2 * C * C * NT;
              t.C = C; t.mode = 3; t.KW = NT; t.S = cs; t.rho = phi;
              if (tp.n == 64) pack_tc(tp, c.st);
            }
            continue;
          }
#endif
          if (!tc_fwd_on(C, cs)) { add(c.W(e, k, i), c.ws + O.pk[i], C, C, 15, 0); continue; }
#ifndef GNAS_EMUL
          tc::TcPackJob& t = tp.j[tp.n++];
          t.src = c.W(e, k, i); t.dst = c.ws + O.pk[i]; t.C = C; t.mode = 0; t.KW = 15;
          if (tp.n == 64) pack_tc(tp, c.st);
#endif
        }
      }
    }
  }
#ifndef GNAS_EMUL
  pack_tc(tp, c.st);
#endif
  if (pb.n) GNAS_LAUNCH(k_pack_weights, dim3(8, pb.n), dim3(256), 0, c.st, pb);
}

void preprocess_forward(Ctx& c) {
  const GnasCellDesc* d = c.d;
  const int C = d->C, B = d->B;
  const long long bs = (long long)C * d->L_p;
  DenseBatch db; db.n = 0; db.B = B; db.Cout = C; db.pad = 0;
  auto job = [&](int i, const float* x, int Cin, int Lin) {
    DenseJob& j = db.j[db.n++];
    j.x = x; j.xbs = (long long)Cin * Lin; j.in_stat = nullptr; j.wp = c.ws + c.P.pre_pk[i];
    j.z = c.ws + c.P.pre_z[i]; j.zbs = bs; j.acc = d->training ? c.acc(i) : nullptr;
    j.Cin = Cin; j.Lin = Lin; j.Lout = d->L_p; j.tin = T_RELU;
  };
  if (d->pre0_factorized) {      // two different kernels on two different inputs: side by side
    job(0, c.s0, d->C_pp, d->L_pp);
    c.rc |= launch_dense_fwd(1, 2, db, c.fork(2));
    job(1, c.s1, d->C_p, d->L_p);
    c.rc |= launch_dense_fwd(1, 1, db, c.st);
    c.join(2);
  } else {
    job(0, c.s0, d->C_pp, d->L_pp);
    job(1, c.s1, d->C_p, d->L_p);
    c.rc |= launch_dense_fwd(1, 1, db, c.st);
  }
  int fs[2] = {0, 1}; float fc[2] = {(float)((long long)B * d->L_p), (float)((long long)B * d->L_p)};
  long long fr[2] = {d->pre_bn_off[0], d->pre_bn_off[1]};
  finalize(c, fs, fc, fr, 2);
  BnApplyBatch ab; ab.n = 2;
  for (int i = 0; i < 2; ++i) {
    ab.j[i].z = c.ws + c.P.pre_z[i]; ab.j[i].y = c.ws + c.P.pre_y[i]; ab.j[i].stat = c.stat(i);
    ab.j[i].gamma = nullptr; ab.j[i].beta = nullptr; ab.j[i].C = C; ab.j[i].L = d->L_p; ab.j[i].total = c.P.nCL_p;
  }
  GNAS_LAUNCH(k_bn_apply, dim3(ew_blocks(c.P.nCL_p), 2), dim3(kThreads), 0, c.st, ab);
}

// -------------------------------------------------------------------------- backward
// temporaries of one edge during a node phase
enum { TMP_SEP15 = 0, TMP_SEP9 = 3, TMP_DIL15 = 6, TMP_DIL9 = 7, TMP_C15 = 8 };  // sep: du2, dyh1, du1

void node_backward(Ctx& c, int node, const float* g, long long gbs, float* const* dsrc, const long long* dsrc_bs, int need_wgrad) {
  const GnasCellDesc* d = c.d;
  const int C = d->C, B = d->B, Lo = d->L_out;
  const long long obs = (long long)C * Lo;
  const float cnt = (float)((long long)B * Lo);
  int edges[5], ne = 0;
  for (int e = 0; e < c.P.n_edges; ++e) if (c.P.e[e].node == node) edges[ne++] = e;
  // ---- R1: sum g, sum g*z for every BN-terminated candidate; direct dmix for identity skips
  {
    RedBatch rb;
    rb.B = B; rb.C = C; rb.L = Lo; rb.nterms = 0; rb.s1_slot = c.P.node_slot[node];
    rb.g = g; rb.gbs = gbs; rb.bacc = reinterpret_cast<double*>(c.scratch + c.P.bacc); rb.dmix = c.dmix();
    for (int i = 0; i < ne; ++i) {
      const int e = edges[i];
      for (int k = 1; k < GNAS_N_PRIMS; ++k) {
        const OpPlan& O = c.P.e[e].op[k];
        if (!O.on) continue;
        RedTerm& t = rb.t[rb.nterms++];
        const int ft = final_tensor(k, c.P.e[e].stride);
        t.widx = c.widx(e, k); t.mask = nullptr;
        if (ft >= 0) { t.z = c.T(e, k, ft); t.bs = obs; t.slot = O.slot[final_slot_idx(k)]; }
        else { c.src(e, t.z, t.bs); t.slot = -1; t.mask = c.mask(e); }
      }
    }
    GNAS_LAUNCH(k_bwd_reduce, dim3(C, cdiv(B, 8), red_term_chunks(rb.nterms)), dim3(kThreads), 0, c.st, rb);
  }
  auto coef_launch = [&](const int* slots, const int* widx, int n, int s1_slot) {
    if (n == 0) return;
    CoefBatch cb;
    cb.bacc = reinterpret_cast<const double*>(c.scratch + c.P.bacc); cb.stat = c.ws + c.P.stat; cb.coef = c.scratch + c.P.coef;
    cb.mixw = c.mixw; cb.dmix = c.dmix(); cb.gamma = nullptr; cb.dgamma = nullptr; cb.dbeta = nullptr;
    cb.C = C; cb.nslots = n; cb.s1_slot = s1_slot;
    for (int i = 0; i < n; ++i) { cb.slot[i] = (short)slots[i]; cb.widx[i] = (short)widx[i]; cb.count[i] = cnt; }
    GNAS_LAUNCH(k_bn_bwd_coef, dim3(n), dim3(C < 256 ? (C + 31) / 32 * 32 : 256), 0, c.st, cb);
  };
  {
    int sl[64], wi[64], n = 0;
    for (int i = 0; i < ne; ++i) {
      const int e = edges[i];
      for (int k = 1; k < GNAS_N_PRIMS; ++k) {
        const OpPlan& O = c.P.e[e].op[k];
        if (!O.on || final_tensor(k, c.P.e[e].stride) < 0) continue;
        sl[n] = O.slot[final_slot_idx(k)]; wi[n] = c.widx(e, k); ++n;
      }
    }
    coef_launch(sl, wi, n, c.P.node_slot[node]);
  }
  const int dwk[4] = {GNAS_SEP15, GNAS_SEP9, GNAS_DIL15, GNAS_DIL9};
  const int dwK[4] = {15, 9, 15, 9}, dwD[4] = {1, 1, 2, 2}, dwP[4] = {7, 4, 14, 8};
  const int tmp0[4] = {TMP_SEP15, TMP_SEP9, TMP_DIL15, TMP_DIL9};
  // ---- pools + identity skip: they only need the coefficients above; their read-modify-write of dX runs on side
  // stream 2 beside the pointwise / depthwise / conv_15 kernels that do not touch dX (joined before the fused edge kernel)
  {
    cudaStream_t pst = c.fork(2);
    PoolBwdBatch pb; pb.n = 0; pb.B = B; pb.C = C; pb.mixw = c.mixw;
    for (int i = 0; i < ne; ++i) {
      const int e = edges[i];
      const EdgePlan& E = c.P.e[e];
      const OpPlan& mx = E.op[GNAS_MAX_POOL]; const OpPlan& av = E.op[GNAS_AVG_POOL]; const OpPlan& sk = E.op[GNAS_SKIP];
      const bool idskip = sk.on && E.stride == 1;
      if (!mx.on && !av.on && !idskip) continue;
      if (pb.n > 0 && (pb.S != E.stride || pb.n == kMaxPool)) launch_pool_bwd(pb, pst);
      pb.S = E.stride;
      PoolBwdJob& j = pb.j[pb.n++];
      j.g = g; j.gbs = gbs; c.src(e, j.x, j.xbs);
      j.zmax = mx.on ? c.T(e, GNAS_MAX_POOL, 0) : nullptr; j.zavg = av.on ? c.T(e, GNAS_AVG_POOL, 0) : nullptr;
      j.coef_max = mx.on ? c.coef(mx.slot[0]) : nullptr; j.coef_avg = av.on ? c.coef(av.slot[0]) : nullptr;
      j.skip_mask = idskip ? c.mask(e) : nullptr;
      j.dx = dsrc[i]; j.dxbs = dsrc_bs[i]; j.Lin = E.Lin; j.Lo = Lo; j.skip_widx = idskip ? c.widx(e, GNAS_SKIP) : -1;
    }
    launch_pool_bwd(pb, pst);
  }
  // The remaining candidates are processed per stride class (edges 0/1 of a reduction cell vs the rest).
  for (int pass = 0; pass < 2; ++pass) {
    int es[5], is[5], n2 = 0;
    for (int i = 0; i < ne; ++i) {
      const bool strided = c.P.e[edges[i]].stride > 1;
      if ((pass == 0) == strided) { es[n2] = edges[i]; is[n2] = i; ++n2; }
    }
    if (n2 == 0) continue;
    const int S = c.P.e[es[0]].stride, Lin = c.P.e[es[0]].Lin;
    auto tmp = [&](int i, int which) { return c.scratch + c.P.tmp[is[i]][which]; };
    cudaStream_t s3 = c.fork(3);      // data gradient of the second conv_15 convolution: independent of the sep / dil chain
    DxBatch xb; xb.n = 0; xb.B = B; xb.Cin = C; xb.pad = 0;
    WgBatch gb; gb.n = 0; gb.B = B; gb.Cout = C;
    auto pw_dx = [&](DxBatch& b) { return tc_pw_on(C) ? launch_conv_dx_tc(b, 1, c.st) : launch_dense_dx(1, 1, b, c.st); };
    auto conv_wg = [&](WgBatch& b, int stride) {
#ifndef GNAS_EMUL
      if (b.n == 0) return 0;
      cudaStream_t ws_ = c.wg_stream();
      if (tc_enabled() && tc_part(stride == 1 ? 5 : 8) && tc::conv_wgrad_tc_supported(b, stride)) return tc::launch_conv_wgrad_tc(b, stride, ws_);
      return launch_conv_wgrad(15, stride, 7, b, ws_);
#endif
      return launch_conv_wgrad(15, stride, 7, b, c.st);
    };
    auto pw_wg = [&](WgBatch& b) {
#ifndef GNAS_EMUL
      if (b.n == 0) return 0;
      cudaStream_t ws_ = c.wg_stream();
      if (tc_enabled() && tc_part(4) && tc::pw_wgrad_tc_supported(b)) return tc::launch_pw_wgrad_tc(b, ws_);
      return launch_pw_wgrad(1, b, ws_);
#endif
      return launch_pw_wgrad(1, b, c.st);
    };
    // -- strided skip (FactorizedReduce): dx + dW
    if (S > 1) {
      for (int i = 0; i < n2; ++i) {
        const int e = es[i];
        const OpPlan& O = c.P.e[e].op[GNAS_SKIP];
        if (!O.on) continue;
        DxJob& j = xb.j[xb.n++];
        j.g = g; j.gbs = gbs; j.z = c.T(e, GNAS_SKIP, 0); j.zbs = obs; j.coef = c.coef(O.slot[0]);
        j.wt = c.W(e, GNAS_SKIP, 0); j.out = dsrc[is[i]]; j.obs = dsrc_bs[is[i]];
        c.src(e, j.xin, j.xibs); j.xin_stat = nullptr; j.bacc = nullptr;
        j.Cout = C; j.Lo = Lo; j.Lin = Lin; j.omode = O_RELU_ACC;
        if (need_wgrad) {
          WgJob& w = gb.j[gb.n++];
          w.g = g; w.gbs = gbs; w.z = j.z; w.zbs = obs; w.coef = j.coef;
          c.src(e, w.r, w.rbs); w.r_stat = nullptr; w.dw = c.dW(e, GNAS_SKIP, 0);
          w.Cin = C; w.Lo = Lo; w.Lin = Lin; w.tin = T_RELU;
        }
      }
      if (xb.n) c.rc |= launch_dense_dx(1, S, xb, c.fork(2));
      if (gb.n) c.rc |= launch_pw_wgrad(S, gb, c.wg_stream());
    }
    // -- last pointwise conv of sep (stage 2) and dil: du = W^T dz, dWpw
    for (int i = 0; i < n2; ++i) {
      const int e = es[i];
      for (int v = 0; v < 4; ++v) {
        const OpPlan& O = c.P.e[e].op[dwk[v]];
        if (!O.on) continue;
        const bool sep = v < 2;
        DxJob& j = xb.j[xb.n++];
        j.g = g; j.gbs = gbs; j.z = c.T(e, dwk[v], sep ? 3 : 1); j.zbs = obs; j.coef = c.coef(O.slot[sep ? 1 : 0]);
        j.wt = tc_pw_on(C) ? c.scratch + O.pkb[sep ? 1 : 0] : c.W(e, dwk[v], sep ? 3 : 1); j.out = tmp(i, tmp0[v]); j.obs = obs;
        j.xin = nullptr; j.xibs = 0; j.xin_stat = nullptr; j.bacc = nullptr;
        j.Cout = C; j.Lo = Lo; j.Lin = Lo; j.omode = O_STORE;
        if (need_wgrad) {
          WgJob& w = gb.j[gb.n++];
          w.g = g; w.gbs = gbs; w.z = j.z; w.zbs = obs; w.coef = j.coef;
          w.r = c.T(e, dwk[v], sep ? 2 : 0); w.rbs = obs; w.r_stat = nullptr; w.dw = c.dW(e, dwk[v], sep ? 3 : 1);
          w.Cin = C; w.Lo = Lo; w.Lin = Lo; w.tin = T_NONE;
        }
        if (xb.n == kMaxDense) { c.rc |= pw_dx(xb); c.rc |= pw_wg(gb); }
      }
    }
    c.rc |= pw_dx(xb);
    c.rc |= pw_wg(gb);
    // -- depthwise: sep stage 2 (-> dyh1 + inner BN sums), both kernel sizes in one launch; the dil depthwise convs
    // (-> dX) join the fused edge kernel below
    {
      DwBwdBatch wb; wb.n = 0; wb.B = B; wb.C = C; wb.pad = 0;
      for (int v = 0; v < 2; ++v)
        for (int i = 0; i < n2; ++i) {
          const int e = es[i];
          const OpPlan& O = c.P.e[e].op[dwk[v]];
          if (!O.on) continue;
          DwBwdJob& j = wb.j[wb.n++];
          j.du = tmp(i, tmp0[v]); j.dubs = obs;
          j.x = c.T(e, dwk[v], 1); j.xbs = obs; j.in_stat = c.stat(O.slot[0]); j.wd = c.W(e, dwk[v], 2);
          j.out = tmp(i, tmp0[v] + 1); j.obs = obs; j.bacc = c.bacc(O.slot[0]);
          j.dwd = need_wgrad ? c.dW(e, dwk[v], 2) : nullptr; j.Lin = Lo; j.Lo = Lo; j.tin = T_BN_RELU; j.omode = O_BNRELU_STORE;
          j.K = dwK[v];
        }
      c.rc |= launch_dw2_bwd(wb, c.st);
    }
    // -- conv_15 second conv: dyh1 + inner BN sums, dW2
    xb.pad = 7;
    for (int i = 0; i < n2; ++i) {
      const int e = es[i];
      const OpPlan& O = c.P.e[e].op[GNAS_CONV15];
      if (!O.on) continue;
      DxJob& j = xb.j[xb.n++];
      j.g = g; j.gbs = gbs; j.z = c.T(e, GNAS_CONV15, 1); j.zbs = obs; j.coef = c.coef(O.slot[1]);
      j.wt = c.scratch + O.pkb[1]; j.out = tmp(i, TMP_C15); j.obs = obs;
      j.xin = c.T(e, GNAS_CONV15, 0); j.xibs = obs; j.xin_stat = c.stat(O.slot[0]); j.bacc = c.bacc(O.slot[0]);
      j.Cout = C; j.Lo = Lo; j.Lin = Lo; j.omode = O_BNRELU_STORE;
      if (need_wgrad) {
        WgJob& w = gb.j[gb.n++];
        w.g = g; w.gbs = gbs; w.z = j.z; w.zbs = obs; w.coef = j.coef;
        w.r = c.T(e, GNAS_CONV15, 0); w.rbs = obs; w.r_stat = c.stat(O.slot[0]); w.dw = c.dW(e, GNAS_CONV15, 1);
        w.Cin = C; w.Lo = Lo; w.Lin = Lo; w.tin = T_BN_RELU;
      }
    }
    c.rc |= tc_dx_on(C, 1, 1) ? launch_conv_dx_tc(xb, 15, s3) : launch_dense_dx(15, 1, xb, s3);
    c.join(3);
    c.rc |= conv_wg(gb, 1);
    // -- inner BN coefficients (w = 1)
    {
      int sl[16], wi[16], n = 0;
      for (int i = 0; i < n2; ++i) {
        const int e = es[i];
        const int ks[3] = {GNAS_SEP15, GNAS_SEP9, GNAS_CONV15};
        for (int t = 0; t < 3; ++t) if (c.P.e[e].op[ks[t]].on) { sl[n] = c.P.e[e].op[ks[t]].slot[0]; wi[n] = -1; ++n; }
      }
      coef_launch(sl, wi, n, -1);
    }
    // -- sep stage 1: pointwise (du1, dWpw1) then depthwise (dX, dWdw1)
    xb.pad = 0;
    for (int i = 0; i < n2; ++i) {
      const int e = es[i];
      for (int v = 0; v < 2; ++v) {
        const OpPlan& O = c.P.e[e].op[dwk[v]];
        if (!O.on) continue;
        DxJob& j = xb.j[xb.n++];
        j.g = tmp(i, tmp0[v] + 1); j.gbs = obs; j.z = c.T(e, dwk[v], 1); j.zbs = obs; j.coef = c.coef(O.slot[0]);
        j.wt = tc_pw_on(C) ? c.scratch + O.pkb[0] : c.W(e, dwk[v], 1); j.out = tmp(i, tmp0[v] + 2); j.obs = obs;
        j.xin = nullptr; j.xibs = 0; j.xin_stat = nullptr; j.bacc = nullptr;
        j.Cout = C; j.Lo = Lo; j.Lin = Lo; j.omode = O_STORE;
        if (need_wgrad) {
          WgJob& w = gb.j[gb.n++];
          w.g = j.g; w.gbs = obs; w.z = j.z; w.zbs = obs; w.coef = j.coef;
          w.r = c.T(e, dwk[v], 0); w.rbs = obs; w.r_stat = nullptr; w.dw = c.dW(e, dwk[v], 1);
          w.Cin = C; w.Lo = Lo; w.Lin = Lo; w.tin = T_NONE;
        }
      }
    }
    c.rc |= pw_dx(xb);
    c.rc |= pw_wg(gb);
    // fused: every depthwise candidate that writes dX of the edge (sep stage 1: du1, dil: du) in one pass over x
    c.join(2);      // pool / skip gradients have landed in dX
    {
      EdgeBwdBatch eb; eb.n = 0; eb.B = B; eb.C = C;
      for (int i = 0; i < n2; ++i) {
        const int e = es[i];
        bool any = false;
        for (int v = 0; v < 4; ++v) any = any || c.P.e[e].op[dwk[v]].on;
        if (!any) continue;
        EdgeBwdJob& j = eb.j[eb.n++];
        for (int v = 0; v < 4; ++v) {
          const bool on = c.P.e[e].op[dwk[v]].on;
          j.du[v] = on ? tmp(i, tmp0[v] + (v < 2 ? 2 : 0)) : nullptr;
          j.wd[v] = on ? c.W(e, dwk[v], 0) : nullptr;
          j.dwd[v] = (on && need_wgrad) ? c.dW(e, dwk[v], 0) : nullptr;
        }
        j.dubs = obs; c.src(e, j.x, j.xbs); j.out = dsrc[is[i]]; j.obs = dsrc_bs[is[i]]; j.Lin = Lin; j.Lo = Lo;
        if (eb.n == kMaxEdgeJobs) c.rc |= launch_edge_bwd(S, eb, c.st);
      }
      c.rc |= launch_edge_bwd(S, eb, c.st);
    }
    // -- conv_15 first conv: dX, dW1
    xb.pad = 7;
    for (int i = 0; i < n2; ++i) {
      const int e = es[i];
      const OpPlan& O = c.P.e[e].op[GNAS_CONV15];
      if (!O.on) continue;
      DxJob& j = xb.j[xb.n++];
      j.g = tmp(i, TMP_C15); j.gbs = obs; j.z = c.T(e, GNAS_CONV15, 0); j.zbs = obs; j.coef = c.coef(O.slot[0]);
      j.wt = c.scratch + O.pkb[0]; j.out = dsrc[is[i]]; j.obs = dsrc_bs[is[i]];
      c.src(e, j.xin, j.xibs); j.xin_stat = nullptr; j.bacc = nullptr;
      j.Cout = C; j.Lo = Lo; j.Lin = Lin; j.omode = O_RELU_ACC;
      if (need_wgrad) {
        WgJob& w = gb.j[gb.n++];
        w.g = j.g; w.gbs = obs; w.z = j.z; w.zbs = obs; w.coef = j.coef;
        c.src(e, w.r, w.rbs); w.r_stat = nullptr; w.dw = c.dW(e, GNAS_CONV15, 0);
        w.Cin = C; w.Lo = Lo; w.Lin = Lin; w.tin = T_RELU;
      }
    }
    c.rc |= tc_dx_on(C, S, 0) ? launch_conv_dx_tc(xb, 15, c.st)
            : (S > 1 && tc_sdx_on(C, S)) ? launch_conv_dx_strided_tc(xb, S, c.st) : launch_dense_dx(15, S, xb, c.st);
    c.rc |= conv_wg(gb, S);
  }
  c.join(1);      // the next node reuses the edge temporaries the forked weight-gradient kernels read
  c.join(2);
  c.join(3);
}

void preprocess_backward(Ctx& c, float* ds0, float* ds1, int need_wgrad) {
  const GnasCellDesc* d = c.d;
  const int C = d->C, B = d->B, Lp = d->L_p;
  const long long bs = (long long)C * Lp;
  const float cnt = (float)((long long)B * Lp);
  for (int i = 0; i < 2; ++i) {
    RedBatch rb;
    rb.B = B; rb.C = C; rb.L = Lp; rb.nterms = 1; rb.s1_slot = -1;
    rb.g = c.scratch + c.P.dS[i]; rb.gbs = bs; rb.bacc = reinterpret_cast<double*>(c.scratch + c.P.bacc); rb.dmix = c.dmix();
    rb.t[0].z = c.ws + c.P.pre_z[i]; rb.t[0].bs = bs; rb.t[0].mask = nullptr; rb.t[0].slot = i; rb.t[0].widx = -1;
    // the two preprocess branches are independent up to the coefficient kernel and again after it: branch 0 on side stream 2
    GNAS_LAUNCH(k_bwd_reduce, dim3(C, cdiv(B, 8), red_term_chunks(rb.nterms)), dim3(kThreads), 0, i == 0 ? c.fork(2) : c.st, rb);
  }
  c.join(2);
  {
    CoefBatch cb;
    cb.bacc = reinterpret_cast<const double*>(c.scratch + c.P.bacc); cb.stat = c.ws + c.P.stat; cb.coef = c.scratch + c.P.coef;
    cb.mixw = c.mixw; cb.dmix = c.dmix(); cb.gamma = nullptr; cb.dgamma = nullptr; cb.dbeta = nullptr;
    cb.C = C; cb.nslots = 2; cb.s1_slot = -1;
    for (int i = 0; i < 2; ++i) { cb.slot[i] = (short)i; cb.widx[i] = -1; cb.count[i] = cnt; }
    GNAS_LAUNCH(k_bn_bwd_coef, dim3(2), dim3(C < 256 ? (C + 31) / 32 * 32 : 256), 0, c.st, cb);
  }
  const float* xs[2] = {c.s0, c.s1};
  float* dxs[2] = {ds0, ds1};
  const int Cin[2] = {d->C_pp, d->C_p}, Lin[2] = {d->L_pp, d->L_p};
  const int S[2] = {d->pre0_factorized ? 2 : 1, 1};
  for (int i = 0; i < 2; ++i) {
    cudaStream_t bst = i == 0 ? c.fork(2) : c.st;
    cudaMemsetAsync(dxs[i], 0, sizeof(float) * (size_t)B * Cin[i] * Lin[i], bst);
    DxBatch xb; xb.n = 1; xb.B = B; xb.Cin = Cin[i]; xb.pad = 0;
    DxJob& j = xb.j[0];
    j.g = c.scratch + c.P.dS[i]; j.gbs = bs; j.z = c.ws + c.P.pre_z[i]; j.zbs = bs; j.coef = c.coef(i);
    j.wt = c.params + d->pre_w_off[i]; j.out = dxs[i]; j.obs = (long long)Cin[i] * Lin[i];
    j.xin = xs[i]; j.xibs = j.obs; j.xin_stat = nullptr; j.bacc = nullptr;
    j.Cout = C; j.Lo = Lp; j.Lin = Lin[i]; j.omode = O_RELU_ACC;
    c.rc |= launch_dense_dx(1, S[i], xb, bst);
    if (need_wgrad) {
      WgBatch gb; gb.n = 1; gb.B = B; gb.Cout = C;
      WgJob& w = gb.j[0];
      w.g = j.g; w.gbs = bs; w.z = j.z; w.zbs = bs; w.coef = j.coef;
      w.r = xs[i]; w.rbs = j.obs; w.r_stat = nullptr; w.dw = c.grads + d->pre_w_off[i];
      w.Cin = Cin[i]; w.Lo = Lp; w.Lin = Lin[i]; w.tin = T_RELU;
      c.rc |= launch_pw_wgrad(S[i], gb, i == 0 ? bst : c.wg_stream());
    }
  }
  c.join(1);
  c.join(2);
}

void pack_backward(Ctx& c) {
  const int C = c.d->C;
  PackBatch pb; pb.n = 0;
#ifndef GNAS_EMUL
  tc::TcPackBatch tp; tp.n = 0;
#endif
  for (int e = 0; e < c.P.n_edges; ++e) {
    const OpPlan& O = c.P.e[e].op[GNAS_CONV15];
    if (!O.on) continue;
    for (int i = 0; i < 2; ++i) {
      const int cs = i == 0 ? c.P.e[e].stride : 1;
#ifndef GNAS_EMUL
      if (i == 0 && cs > 1 && tc_sdx_on(C, cs)) {
        const int NT = strided_dx_taps(cs);
        for (int rho = 0; rho < cs; ++rho) {
          tc::TcPackJob& t = tp.j[tp.n++];
          t.src = c.W(e, GNAS_CONV15, 0); t.dst = c.scratch + O.pkb[0] + (long long)rho * 2 * C * C * NT;
          t.C = C; t.mode = 2; t.KW = NT; t.S = cs; t.rho = rho;
          if (tp.n == 64) pack_tc(tp, c.st);
        }
        continue;
      }
#endif
      if (tc_dx_on(C, cs, i)) {
#ifndef GNAS_EMUL
        tc::TcPackJob& t = tp.j[tp.n++];
        t.src = c.W(e, GNAS_CONV15, i); t.dst = c.scratch + O.pkb[i]; t.C = C; t.mode = 1; t.KW = 15;
        if (tp.n == 64) pack_tc(tp, c.st);
#endif
        continue;
      }
      PackJob& j = pb.j[pb.n++];
      j.src = c.W(e, GNAS_CONV15, i); j.dst = c.scratch + O.pkb[i]; j.Cout = C; j.Cin = C; j.KW = 15; j.mode = 1;
    }
  }
#ifndef GNAS_EMUL
  if (tc_pw_on(C)) {
    const int ks[4] = {GNAS_SEP15, GNAS_SEP9, GNAS_DIL15, GNAS_DIL9};
    for (int e = 0; e < c.P.n_edges; ++e)
      for (int v = 0; v < 4; ++v) {
        const OpPlan& O = c.P.e[e].op[ks[v]];
        if (!O.on) continue;
        for (int i = 0; i < (v < 2 ? 2 : 1); ++i) {
          tc::TcPackJob& t = tp.j[tp.n++];
          t.src = c.W(e, ks[v], i == 0 ? 1 : 3); t.dst = c.scratch + O.pkb[i]; t.C = C; t.mode = 1; t.KW = 1;
          if (tp.n == 64) pack_tc(tp, c.st);
        }
      }
  }
  pack_tc(tp, c.st);
#endif
  if (pb.n) GNAS_LAUNCH(k_pack_weights, dim3(8, pb.n), dim3(256), 0, c.st, pb);
}

}  // namespace
}  // namespace gnas

using namespace gnas;

extern "C" int gnas_version(void) { return 100; }
extern "C" const char* gnas_last_error(void) { return gnas::get_error(); }

extern "C" int gnas_cell_workspace(const GnasCellDesc* d, int64_t* fwd_floats, int64_t* bwd_floats) {
  int rc = validate(d);
  if (rc) return rc;
  CellPlan P;
  make_plan(d, P);
  if (fwd_floats) *fwd_floats = P.fwd_total;
  if (bwd_floats) *bwd_floats = P.bwd_total;
  return 0;
}

extern "C" int gnas_cell_forward(const GnasCellDesc* d, const float* params, float* bn_running,
                                 const float* s0, const float* s1, const float* mixw,
                                 const int64_t* skip_mask_host, float* out, float* ws, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if (!params || !s1 || !mixw || !out || !ws || (d->has_pre && !s0)) { set_error("cell_forward: null pointer"); return GNAS_ERR_ARG; }
  Ctx c;
  c.d = d; make_plan(d, c.P);
  c.params = params; c.grads = nullptr; c.running = bn_running; c.s0 = s0; c.s1 = s1; c.mixw = mixw;
  c.skip_mask = skip_mask_host; c.out = out; c.ws = ws; c.scratch = nullptr; c.st = (cudaStream_t)stream;
  c.init_sides(true);
  if (d->training)
    cudaMemsetAsync(ws + c.P.acc, 0, sizeof(double) * (size_t)c.P.nslots * 2 * d->C, c.st);
  pack_forward(c);
  if (d->has_pre) {
    preprocess_forward(c);
    const int ph0[8] = {0, 1, 2, 3, 5, 6, 9, 10};   // edges reading s0'/s1'
    edges_forward(c, ph0, 8);
    node_mix(c, 0);
    const int ph1[3] = {4, 7, 11};                  // edges reading node 0
    edges_forward(c, ph1, 3);
    node_mix(c, 1);
    const int ph2[2] = {8, 12};                     // edges reading node 1
    edges_forward(c, ph2, 2);
    node_mix(c, 2);
    const int ph3[1] = {13};                        // edge reading node 2
    edges_forward(c, ph3, 1);
    node_mix(c, 3);
  } else {
    const int e0[1] = {0};
    edges_forward(c, e0, 1);
    node_mix(c, 0);
  }
  if (c.rc) return c.rc;
  return check_launch("gnas_cell_forward");
}

extern "C" int gnas_cell_backward(const GnasCellDesc* d, const float* params, float* grads,
                                  const float* s0, const float* s1, const float* mixw,
                                  const int64_t* skip_mask_host, const float* out, const float* dout,
                                  const float* ws, float* scratch, float* ds0, float* ds1, float* dmixw,
                                  int need_wgrad, void* stream) {
  int rc = validate(d);
  if (rc) return rc;
  if (!params || !s1 || !mixw || !out || !dout || !ws || !scratch || !ds1 || !dmixw || (need_wgrad && !grads) ||
      (d->has_pre && (!s0 || !ds0))) {
    set_error("cell_backward: null pointer"); return GNAS_ERR_ARG;
  }
  if (!d->training) {   // the kernels differentiate batch-statistics BatchNorm only
    set_error("cell_backward: the forward ran in eval mode (running statistics); its backward is not implemented");
    return GNAS_ERR_UNSUPPORTED;
  }
  Ctx c;
  c.d = d; make_plan(d, c.P);
  c.params = params; c.grads = grads; c.running = nullptr; c.s0 = s0; c.s1 = s1; c.mixw = mixw;
  c.skip_mask = skip_mask_host; c.out = const_cast<float*>(out); c.ws = const_cast<float*>(ws); c.scratch = scratch;
  c.st = (cudaStream_t)stream;
  c.init_sides(true);
  const int C = d->C, B = d->B, Lo = d->L_out;
  const long long obs = (long long)C * Lo;
  // bacc | dmix | coef sit at the head of the scratch: zero the accumulators
  cudaMemsetAsync(scratch, 0, sizeof(float) * (size_t)c.P.coef, c.st);
  pack_backward(c);
  if (d->has_pre) {
    for (int i = 0; i < 4; ++i) {
      SliceCopy sc;
      sc.src = dout + (long long)i * obs; sc.sbs = 4 * obs; sc.dst = scratch + c.P.G[i]; sc.dbs = obs; sc.B = B; sc.CL = (int)obs;
      GNAS_LAUNCH(k_copy_rows, dim3(ew_blocks((long long)B * obs)), dim3(kThreads), 0, c.st, sc);
    }
    cudaMemsetAsync(scratch + c.P.dS[0], 0, sizeof(float) * (size_t)(c.P.dS[1] - c.P.dS[0]) * 2, c.st);
    for (int node = 3; node >= 0; --node) {
      float* dsrc[5]; long long dbs[5];
      for (int j = 0; j < node + 2; ++j) {
        if (j < 2) { dsrc[j] = scratch + c.P.dS[j]; dbs[j] = (long long)C * d->L_p; }
        else { dsrc[j] = scratch + c.P.G[j - 2]; dbs[j] = obs; }
      }
      node_backward(c, node, scratch + c.P.G[node], obs, dsrc, dbs, need_wgrad);
    }
    preprocess_backward(c, ds0, ds1, need_wgrad);
  } else {
    cudaMemsetAsync(ds1, 0, sizeof(float) * (size_t)B * C * d->L_p, c.st);
    float* dsrc[1] = {ds1}; long long dbs[1] = {(long long)C * d->L_p};
    node_backward(c, 0, dout, obs, dsrc, dbs, need_wgrad);
  }
  const int nm = d->n_edges * d->n_prims;
  GNAS_LAUNCH(k_double_to_float, dim3(cdiv(nm, 128)), dim3(128), 0, c.st, (const double*)c.dmix(), dmixw, nm, 0);
  if (c.rc) return c.rc;
  return check_launch("gnas_cell_backward");
}

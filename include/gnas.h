/* gnas.h -- C ABI of the B200-native genomeDARTS supernet search step.
 *
 * The reference (GenomeNet/Genome-NAS) has no FFI layer: its hot path is Python
 * nn.Modules whose arithmetic runs inside ATen/cuDNN.  This header is the
 * boundary a maintainer binds (ctypes stub in INTEGRATION.md) to replace that
 * arithmetic; each entry point names the reference interface it stands in for.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *     descriptors (GnasCellDesc, GnasRhnDesc) are read on the host.
 *   - all tensors are contiguous fp32, layouts as in the reference
 *     (activations [B][C][L], RHN sequences [T][B][H]).
 *   - return value: 0 ok, <0 argument/shape error, >0 cudaError_t.  Nothing
 *     throws, exits or allocates; workspaces are caller-owned (sizes from the
 *     *_workspace queries); gnas_last_error() describes the last failure.
 *   - `stream` is a cudaStream_t passed as void*.  Entry points are
 *     asynchronous and re-entrant across streams/devices.
 */
#ifndef GNAS_H_
#define GNAS_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNAS_MAX_EDGES 14
#define GNAS_N_PRIMS 9
#define GNAS_RNN_STEPS 8
#define GNAS_RNN_EDGES 36
#define GNAS_RNN_PRIMS 5

/* primitive order = generalNAS_tools/genotypes.py:29-39 (PRIMITIVES_cnn) */
enum { GNAS_NONE = 0, GNAS_MAX_POOL = 1, GNAS_AVG_POOL = 2, GNAS_SKIP = 3, GNAS_SEP15 = 4,
       GNAS_SEP9 = 5, GNAS_DIL15 = 6, GNAS_DIL9 = 7, GNAS_CONV15 = 8 };

int gnas_version(void);
const char* gnas_last_error(void);
/* number of kernels this library has launched in this process (monotonic) */
int64_t gnas_kernel_launches(void);
/* per-kernel device timing for bench.py: while enabled every launch is bracketed by CUDA events on its
 * stream; gnas_profile_report synchronises, writes "kernel<TAB>count<TAB>total_ms" lines into buf
 * (returns bytes written, <0 if cap is too small) and clears the records. */
int gnas_profile_enable(int on);
int64_t gnas_profile_report(char* buf_host, int64_t cap);
/* Independent kernel chains of a pass (weight gradients, the conv_15 chain, pool / skip gradients, the off-critical-
 * path products of the RHN backward) run on library-owned side streams, forked from and joined back into the caller's
 * stream by events inside every call (a CUDA-graph capture records them as dependencies).  on = 0 serialises
 * everything on the caller's stream (bench.py does so for its per-kernel timing steps); returns the previous setting. */
int gnas_side_streams(int on);
/* median elapsed time (ms) of n empty event pairs on `stream`: the bracketing overhead, measured live */
double gnas_profile_event_overhead_ms(int n, void* stream);

/* ---- CNN search cell: darts_tools/model_discCNN.py:96-201 (CNN_Cell_search) with its
 * MixedOps (:46-93) and the operator zoo generalNAS_tools/operations_14_9.py:17-300.
 * A MixedOp on its own is the same descriptor with n_edges = 1 and has_pre = 0. */
typedef struct {
  int32_t B;                 /* batch rows */
  int32_t C;                 /* cell channels */
  int32_t C_pp, C_p;         /* channels of s0 / s1 (ignored when has_pre = 0) */
  int32_t L_pp, L_p;         /* lengths of s0 / s1; L_p is the length the edges from s0'/s1' read */
  int32_t L_out;             /* node length = (L_p - 1) / stride + 1 */
  int32_t stride;            /* stride of edges reading the two input states (1, 2 or 3) */
  int32_t n_edges;           /* 14 for a cell (2+3+4+5), 1 for a lone MixedOp */
  int32_t has_pre;           /* 1: preprocess0/1 + 4 nodes + concat; 0: single edge on x = s1 */
  int32_t pre0_factorized;   /* preprocess0 is FactorizedReduce(stride 2) (reduction_prev) */
  int32_t n_prims;           /* enabled primitives per edge = width of the mixing-weight rows */
  int32_t training;          /* 1: batch statistics + running-stat update; 0: running statistics */
  float bn_momentum;         /* 0.1 */
  uint8_t sw[GNAS_MAX_EDGES][GNAS_N_PRIMS];            /* switches */
  /* offsets (in floats) into the flat parameter / gradient buffers, -1 if absent.
   * per primitive: sep_conv {dw1, pw1, dw2, pw2}; dil_conv {dw, pw}; conv_15 {conv1, conv2};
   * strided skip {conv_1 (conv_2 follows contiguously)}. */
  int64_t w_off[GNAS_MAX_EDGES][GNAS_N_PRIMS][4];
  int64_t pre_w_off[2];
  /* offsets into the flat BatchNorm running-stat buffer (mean at off, var at off + C), -1 if absent */
  int64_t bn_off[GNAS_MAX_EDGES][GNAS_N_PRIMS][2];
  int64_t pre_bn_off[2];
} GnasCellDesc;

/* floats needed for the forward workspace (kept until backward) and the backward scratch */
int gnas_cell_workspace(const GnasCellDesc* d, int64_t* fwd_floats, int64_t* bwd_floats);

/* out: [B][4C][L_out] (cell) or [B][C][L_out] (lone MixedOp).  mixw: [n_edges][n_prims] =
 * softmax(alpha) rows.  skip_mask: [n_edges] device pointers-as-int64 on the HOST (0 = none):
 * dropout masks of the identity skips (P-DARTS, model_discCNN.py:60-61). */
int gnas_cell_forward(const GnasCellDesc* d, const float* params, float* bn_running,
                      const float* s0, const float* s1, const float* mixw,
                      const int64_t* skip_mask_host, float* out, float* ws, void* stream);

/* out: the forward output (its node slices are inputs of later edges).  ds0/ds1 are overwritten; grads (same layout as params) and dmixw are accumulated into /
 * overwritten respectively.  need_wgrad = 0 skips every weight-gradient kernel (the first-order
 * alpha step discards them: darts_tools/architect.py:69-72). */
int gnas_cell_backward(const GnasCellDesc* d, const float* params, float* grads,
                       const float* s0, const float* s1, const float* mixw,
                       const int64_t* skip_mask_host, const float* out, const float* dout,
                       const float* ws, float* scratch, float* ds0, float* ds1, float* dmixw,
                       int need_wgrad, void* stream);

/* ---- stem: Conv1d(4, C0, 13, pad 6, no bias) + BatchNorm1d(C0) affine.
 * darts_tools/model_discCNN.py:350-353,493.  ws floats: gnas_stem_workspace. */
int gnas_stem_workspace(int B, int C0, int L, int64_t* fwd_floats, int64_t* bwd_floats);
int gnas_stem_forward(int B, int C0, int L, int training, float momentum, const float* x,
                      const float* w, const float* gamma, const float* beta, float* running,
                      float* out, float* ws, void* stream);
int gnas_stem_backward(int B, int C0, int L, const float* x, const float* w, const float* gamma,
                       const float* dout, const float* ws, float* scratch, float* dw,
                       float* dgamma, float* dbeta, void* stream);

/* ---- compact input: one byte per base (0..3 = hot channel, anything else = all-zero column, e.g. N) expanded on the
 * device into the fp32 one-hot [B][4][L] the stem reads -- 16x fewer host->device bytes than the tensor the reference's
 * loader builds on the host (generalNAS_tools/data_preprocessing_TF.py:90-149). */
int gnas_onehot_expand(const uint8_t* codes, int64_t B, int64_t L, float* x, void* stream);

/* ---- recurrent highway layer of the search cell: darts_tools/model_discCNN.py:226-268
 * (DARTSCell.forward, _compute_init_state) + model_search.py:54-97 (DARTSCellSearch.cell). */
typedef struct {
  int32_t T, B, H;           /* time steps, batch rows, hidden size (ninp == nhid == H) */
  int32_t training;
  int32_t n_prims;           /* enabled activations per edge = columns of the probability matrix */
  float bn_momentum;
  uint8_t sw[GNAS_RNN_EDGES][GNAS_RNN_PRIMS];
} GnasRhnDesc;

int gnas_rhn_workspace(const GnasRhnDesc* d, int64_t* fwd_floats, int64_t* bwd_floats);
/* x [T][B][H]; h0 [B][H]; W0 [2H][2H]; Ws [8][H][2H]; probs [36][n_prims] = softmax(weights);
 * x_mask/h_mask [B][H] (utils.py:354-363) or null in eval; bn_running: mean[H] | var[H];
 * out [T][B][H] (= hiddens; the last row is the returned hidden state). */
int gnas_rhn_forward(const GnasRhnDesc* d, const float* x, const float* h0, const float* W0,
                     const float* Ws, const float* probs, const float* x_mask, const float* h_mask,
                     float* bn_running, float* out, float* ws, void* stream);
int gnas_rhn_backward(const GnasRhnDesc* d, const float* x, const float* h0, const float* W0,
                      const float* Ws, const float* probs, const float* x_mask, const float* h_mask,
                      const float* dout, float* ws, float* scratch, float* dx, float* dh0,
                      float* dW0, float* dWs, float* dprobs, int need_wgrad, void* stream);

/* ---- classification head: permute(1,0,2) + flatten + Dropout + Linear(T*H, D1) + ReLU + Linear(D1, NC)
 * (+ Sigmoid when sigmoid = 1): darts_tools/model_discCNN.py:547-566.
 * hs [T][B][H] (the RHN output); drop_mask [B][T*H] = keep / (1 - p) or null; Wd [D1][T*H], bd [D1],
 * Wc [NC][D1], bc [NC] (bias pointers may be null); out [B][NC].  ws (gnas_head_workspace, 16-byte aligned) is kept
 * until backward.  Backward overwrites dhs [T][B][H] and, when need_wgrad = 1, ACCUMULATES dWd / dbd / dWc / dbc. */
int gnas_head_workspace(int T, int B, int H, int D1, int NC, int64_t* fwd_floats, int64_t* bwd_floats);
int gnas_head_forward(int T, int B, int H, int D1, int NC, int sigmoid, const float* hs,
                      const float* drop_mask, const float* Wd, const float* bd, const float* Wc,
                      const float* bc, float* out, float* ws, void* stream);
int gnas_head_backward(int T, int B, int H, int D1, int NC, int sigmoid, const float* drop_mask,
                       const float* Wd, const float* Wc, const float* out, const float* dout,
                       const float* ws, float* scratch, float* dhs, float* dWd, float* dbd,
                       float* dWc, float* dbc, int need_wgrad, void* stream);

/* ---- loss terms of the search step: generalNAS_tools/train_and_validate.py:136-140.
 * gnas_bce_*: torch.nn.BCELoss(reduction='mean') over n probabilities (logs clamped at -100; backward divides by
 * max(p (1 - p), 1e-12)); gnas_tar_*: mean((h[1:] - h[:-1])^2) over h [T][BH] (the caller scales by beta).
 * loss / gout are DEVICE scalars; ws holds gnas_loss_workspace floats; sums run in double in a fixed order. */
int gnas_loss_workspace(int64_t* floats);
int gnas_bce_forward(const float* p, const float* y, int64_t n, float* loss, float* ws, void* stream);
int gnas_bce_backward(const float* p, const float* y, const float* gout, int64_t n, float* dp, void* stream);
int gnas_tar_forward(const float* h, int T, int64_t BH, float* loss, float* ws, void* stream);
int gnas_tar_backward(const float* h, const float* gout, int T, int64_t BH, float* dh, void* stream);

/* ---- optimiser steps on flat buffers.
 * gnas_grad_sumsq: out[0] += sum g^2 (double) -- global-norm clip, train_and_validate.py:146-147.
 * gnas_sgd_step:   torch.optim.SGD semantics (weight decay, momentum, first-step buffer copy),
 *                  grads scaled by min(1, clip/(sqrt(sumsq)+1e-6)) read from the device.
 * gnas_adam_step:  torch.optim.Adam semantics used by Architect (architect.py:42,66). */
int gnas_grad_sumsq(const float* g, int64_t n, double* out, void* stream);
int gnas_sgd_step(float* p, const float* g, float* mom, int64_t n, float lr, float wd,
                  float momentum, int first_step, const double* sumsq, float clip, void* stream);
/* step_dev (optional, device int): when given it is incremented and used for the bias correction instead of
 * `step`, so the call can be replayed from a CUDA graph.  sumsq (optional, device double = sum g^2 from
 * gnas_grad_sumsq): gradients are scaled by min(1, clip/(sqrt(sumsq)+1e-6)) first -- the P-DARTS inline
 * alpha step clips the architecture gradients (train_and_validate.py:87-101). */
int gnas_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float wd,
                   float beta1, float beta2, float eps, int step, int* step_dev,
                   const double* sumsq, float clip, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GNAS_H_ */

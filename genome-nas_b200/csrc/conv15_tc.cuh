// Dense stride-1 Conv1d C -> C with 15 taps (conv_15) or 1 tap (the pointwise convs of sep_conv / dil_conv,
// operations_14_9.py:36-64) on the 5th-generation tensor cores:
// implicit GEMM  D[128 positions][C_out] += A_tap[128][8 ch] * B_tap[C_out][8 ch]^T  issued as
// tcgen05.mma kind::tf32 with the accumulator in TMEM, error-compensated 3xTF32
// (a_hi*b_hi + a_hi*b_lo + a_lo*b_hi, fp32 accumulate) because plain TF32 destroys the gradients
// (SURVEY.md App. C).  Replaces the arithmetic of generalNAS_tools/operations_14_9.py:67-81 for the
// stride-1 convolutions of conv_15 (forward and, with transposed / tap-flipped weights, the data gradient).
//
// Shared-memory layouts are the no-swizzle K-major canonical form ((8,m),(4,2)):((16B,SBO),(4B,LBO)):
//   A (activations): [4-channel group][row][4 floats]; a row is a sequence position, so the tap shift of the
//                    implicit GEMM is a +16-byte offset of the descriptor start address (no im2col copy).
//   B (weights)    : [tap][4-channel group][C_out row][4 floats], pre-split into hi/lo by k_pack_tc.
#pragma once
#include "common.cuh"
#include "cnn_fwd.cuh"
#include "tc_common.cuh"

#ifndef GNAS_EMUL
namespace gnas {
namespace tc {

// ------------------------------------------------------------------ weight packing (hi | lo, tensor-core layout)
// dst[half][g][tap][kc][n][e]  (g: group of 8 reduction channels, kc: 4-channel core matrix, n: output row, e: 0..3)
// mode 0 (forward): n = co, reduction channel = ci, tap as is;   mode 1 (data gradient): n = ci, reduction = co,
// tap flipped (out[m] = sum_t W[.,.,t] dz[m + pad - t]  ==  conv with taps t' = KW - 1 - t).
struct TcPackJob { const float* src; float* dst; int C, mode, KW, S, rho; };   // mode 2: residue class rho of a stride-S data gradient
struct TcPackBatch { int n; TcPackJob j[64]; };
static __global__ void __launch_bounds__(256) k_pack_tc(const __grid_constant__ TcPackBatch P) {
  const TcPackJob& J = P.j[blockIdx.y];
  const int C = J.C, KW = J.KW, total = C * C * KW;
  // mode 2: data gradient of a stride-S conv_15 restricted to output positions m = S q + rho.  Only the taps
  // t = t0 + S k (t0 = (rho + 7) mod S) reach them; written as a stride-1 correlation over dz with KW taps,
  // tap k' uses t = t0 + S (KW - 1 - k')  (t > 14: zero weight pads the class to KW taps).
  // mode 3: forward of a stride-S conv_15 restricted to the input phase rho (x[S q + rho]): taps t = t0 + S k.
  const int t0 = J.mode >= 2 ? (J.rho + 7) % J.S : 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i % KW, ci = (i / KW) % C, co = i / (KW * C);
    float w;
    int n, r, tap;
    if (J.mode == 2) {
      const int tt = t0 + J.S * (KW - 1 - t);
      w = tt <= 14 ? J.src[((long long)co * C + ci) * 15 + tt] : 0.f;
      n = ci; r = co; tap = t;
    } else if (J.mode == 3) {
      const int tt = t0 + J.S * t;
      w = tt <= 14 ? J.src[((long long)co * C + ci) * 15 + tt] : 0.f;
      n = co; r = ci; tap = t;
    } else {
      w = J.src[i];
      n = J.mode == 0 ? co : ci; r = J.mode == 0 ? ci : co; tap = J.mode == 0 ? t : KW - 1 - t;
    }
    const int g = r >> 3, kc = (r >> 2) & 1, e = r & 3;
    const long long o = ((((long long)g * KW + tap) * 2 + kc) * C + n) * 4 + e;
    const float hi = __uint_as_float(to_tf32(w));
    J.dst[o] = hi;
    J.dst[(long long)total + o] = __uint_as_float(to_tf32(w - hi));
  }
}

#ifdef GNAS_TC_MARKS
__device__ long long g_tc_marks[64];
#define TC_MARK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_tc_marks[i] = clock64(); } while (0)
#else
#define TC_MARK(i) do { } while (0)
#endif

// ------------------------------------------------------------------ the convolution
struct TcJob {
  const float* x; long long xbs;      // input [B][C][L] (forward: x or z1; data gradient: g)
  const float* x2; long long x2bs;    // data gradient: raw BN input z of the conv output (dz = A*g + Bz*z + D), or null
  const float* in_stat;               // forward T_BN_RELU: mean | rstd of the producer BN;  data gradient: A | Bz | D
  const float* w;                     // packed hi | lo (k_pack_tc)
  float* z; long long zbs;            // output [B][C][L]
  double* acc;                        // forward / O_BNRELU_STORE: per-channel (sum, sumsq) or (sum v, sum v*z1)
  const float* mref; long long mrbs;  // epilogue mask reference: x (O_RELU_ACC) or z1 (O_BNRELU_STORE)
  const float* mstat;                 // mean | rstd of the inner BN (O_BNRELU_STORE)
  int L, tin, omode;                  // tin: T_RELU / T_BN_RELU / -1 (affine dz); omode: O_STORE(+stats) / O_RELU_ACC / O_BNRELU_STORE
  int zl, zs, zo;                     // output (and mref) geometry: row length, position stride and offset (L, 1, 0 unless strided)
  int halo;                           // input row of output row 0 is position -halo; < 0: (KW - 1) / 2
  int xl, xs, xo;                     // input geometry: channel pitch, position stride and offset (L, 1, 0 unless decimated)
};
struct TcBatch { int n, B, C, KW; TcJob j[kMaxDense]; };

// How the 128 accumulator rows of a CTA map to (batch row, position).  Long sequences: 128 consecutive positions
// of one batch row.  Short ones (L + KW - 1 <= 64, e.g. the 42-position cell): several batch rows side by side,
// each in a segment of L + KW - 1 rows so that the tap shift never crosses into the neighbour's data.
struct TcTiling {
  int seg, nseg, tiles_per_row;
  __host__ __device__ TcTiling(int L, int KW) {
    seg = L + KW - 1;
    nseg = seg <= 64 ? (128 - L) / seg + 1 : 1;
    tiles_per_row = (L + 127) / 128;
  }
  __host__ __device__ int tiles(int B) const { return nseg > 1 ? (B + nseg - 1) / nseg : B * tiles_per_row; }
};

constexpr int kNAcc = 4;   // TMEM accumulators per tile: 3 take the hi*hi products in turn, 1 the cross terms

// lane i ends with sum over the warp of a[i] (31 shuffles instead of 32 five-step reductions)
__device__ __forceinline__ float warp_transpose_sum(float (&a)[32], int lane) {
#pragma unroll
  for (int lvl = 0; lvl < 5; ++lvl) {
    const int off = 16 >> lvl;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = upper ? a[k] : a[k + off];
      const float keep = upper ? a[k + off] : a[k];
      a[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return a[0];
}

template <int C, int KW> struct TcCfg {
  static constexpr int ROWS = KW == 1 ? 128 : 144;         // A rows per 4-channel group: 128 outputs + KW-1 halo (+2 pad)
  static constexpr int NU = (C / 8) * KW;                  // MMA units: (group of 8 reduction channels, tap)
  // units per weight chunk / weight stages in shared memory (KW = 5, 8: residue classes of the strided data gradient)
  static constexpr int TT = KW == 1 ? (C == 128 ? 4 : C / 8) : KW == 15 ? (C == 32 ? 5 : 3) : (C == 128 ? 1 : (KW == 5 ? 5 : 4));
  static constexpr int NST = KW == 1 ? (C == 128 ? 2 : 1) : KW == 15 ? (C == 32 ? 4 : 3) : (C == 128 ? 8 : 3);
  static constexpr int NCH = NU / TT;
  static constexpr int NH = NU < 3 ? NU : 3;               // accumulators the hi*hi products rotate over
  static constexpr int W_UNIT = 2 * C * 4;                 // floats of one unit (2 core-matrix columns x C rows), hi or lo
  static constexpr int W_CHUNK = TT * W_UNIT;
  static constexpr int A_HALF = (C / 4) * ROWS * 4;        // floats of A_hi (= A_lo)
  static constexpr size_t smem = sizeof(float) * (size_t)(2 * A_HALF + NST * 2 * W_CHUNK + 8 * C) + 128;
  static_assert(NU % TT == 0, "chunks must tile the units");
};

// Warp roles: thread 0 issues the MMAs, thread 32 streams the weight chunks (TMA bulk copies into an NST-deep
// ring, full/empty mbarriers), warps 2-7 build the activation tile group by group (the MMAs of group g start
// as soon as its 8 channels are in shared memory), all 8 warps run the epilogue.
template <int C, int KW>
__global__ void __launch_bounds__(kTcThreads, C == 128 ? 1 : 2) k_conv_tc(const __grid_constant__ TcBatch P) {
  using Cfg = TcCfg<C, KW>;
  constexpr int G = C / 8;                                  // reduction groups of 8 channels
  constexpr int TT = Cfg::TT, NCH = Cfg::NCH, NST = Cfg::NST, ROWS = Cfg::ROWS;
  constexpr int A_HALF = Cfg::A_HALF, W_CHUNK = Cfg::W_CHUNK, W_UNIT = Cfg::W_UNIT;
  constexpr uint32_t LBO_A = ROWS * 16, SBO = 128, LBO_B = C * 16;
  constexpr int TM_COLS = kNAcc * C;
  // activation-tile threads: warps 2..7; for the pointwise kernel (one or few weight chunks, nothing to overlap)
  // the producer and MMA warps help building the tile first
  constexpr int NLD = KW == 1 ? kTcThreads : kTcThreads - 64, LT0 = kTcThreads - NLD;
  const TcJob& J = P.j[blockIdx.z];
  const int L = J.L;
  const int HALO = J.halo >= 0 ? J.halo : (KW - 1) / 2;
  const TcTiling tl(L, KW);
  const bool packed = tl.nseg > 1;
  if ((int)blockIdx.x >= tl.tiles(P.B)) return;
  const int b0 = packed ? blockIdx.x * tl.nseg : blockIdx.x / tl.tiles_per_row;
  const int l0 = packed ? 0 : (blockIdx.x % tl.tiles_per_row) * 128;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler (uniform datapath)
  extern __shared__ __align__(128) unsigned char tc_smem_raw[];
  float* A_hi = reinterpret_cast<float*>(tc_smem_raw);
  float* A_lo = A_hi + A_HALF;
  float* Wb = A_lo + A_HALF;                                // [NST][hi | lo][W_CHUNK]
  double* ssum = reinterpret_cast<double*>(Wb + NST * 2 * W_CHUNK);   // [2][C] in double: order-independent statistics
  float* sp = reinterpret_cast<float*>(ssum + 2 * C);       // [3][C] input-transform parameters
  float* smean = sp + 3 * C;                                // [C] mean of the inner BN (O_BNRELU_STORE)
  __shared__ __align__(8) uint64_t bar_full[NST], bar_empty[NST], bar_a[G], bar_done;
  __shared__ uint32_t tmem_base_s;

  TC_MARK(0);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
#pragma unroll
    for (int i = 0; i < NST; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
#pragma unroll
    for (int i = 0; i < G; ++i) mbar_init(&bar_a[i], NLD);
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 2 * C; i += kTcThreads) ssum[i] = 0.0;
  {
    const int np = J.tin == T_BN_RELU ? 2 * C : (J.tin < 0 && J.x2 ? 3 * C : 0);
    for (int i = tid; i < np; i += kTcThreads) sp[i] = J.in_stat[i];
    if (J.omode == O_BNRELU_STORE) for (int i = tid; i < C; i += kTcThreads) smean[i] = J.mstat[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  TC_MARK(1);
  const uint32_t tmem_d = tmem_base_s;

  auto produce = [&](int c_begin, int c_end) {
    // ---- weight producer: chunk c = (group g, tap group) is W_CHUNK contiguous floats in each of the hi / lo halves
    const float* whi = J.w;
    const float* wlo = J.w + (long long)C * C * KW;
    for (int c = c_begin; c < c_end; ++c) {
      const int s = c % NST;
      if (c >= NST) mbar_wait(&bar_empty[s], ((c / NST) - 1) & 1);
      if (elect_one()) {
        float* dst = Wb + s * 2 * W_CHUNK;
        mbar_expect_tx(&bar_full[s], 2 * W_CHUNK * 4);
        bulk_g2s(dst, whi + (long long)c * W_CHUNK, W_CHUNK * 4, &bar_full[s]);
        bulk_g2s(dst + W_CHUNK, wlo + (long long)c * W_CHUNK, W_CHUNK * 4, &bar_full[s]);
      }
      __syncwarp();
    }
  };
  auto mma_loop = [&]() {
    // ---- MMA issuer: the whole warp runs the loop (descriptors stay in uniform registers), one elected lane issues
    const uint32_t idesc = make_idesc(C);
    const uint64_t ah0 = make_desc(smem_u32(A_hi), LBO_A, SBO), al0 = make_desc(smem_u32(A_lo), LBO_A, SBO);
    const uint64_t bh0 = make_desc(smem_u32(Wb), LBO_B, SBO);
    int gw = 0;                    // activation groups already waited for
    for (int c = 0; c < NCH; ++c) {
      const int s = c % NST;
      const int g_last = (c * TT + TT - 1) / KW;
      while (gw <= g_last) { mbar_wait(&bar_a[gw], 0); ++gw; }
      if (c == 0) TC_MARK(2);
      mbar_wait(&bar_full[s], (c / NST) & 1);
      tc_fence_after();
      if (elect_one()) {
        // descriptor start-address field is (byte address >> 4) in the low 14 bits: offsets are plain adds
        const uint64_t bhc = bh0 + (uint64_t)(((uint32_t)s * 2 * W_CHUNK * 4) >> 4), blc = bhc + (uint64_t)((W_CHUNK * 4) >> 4);
#pragma unroll
        for (int t = 0; t < TT; ++t) {
          const int unit = c * TT + t, g = unit / KW, tap = unit - g * KW;
          const uint64_t a_off = (uint64_t)(((uint32_t)(g * 2) * LBO_A + (uint32_t)tap * 16) >> 4);   // tap shift = +16 B per row
          const uint64_t ah = ah0 + a_off, al = al0 + a_off;
          const uint64_t bh = bhc + (uint64_t)((t * W_UNIT * 4) >> 4), bl = blc + (uint64_t)((t * W_UNIT * 4) >> 4);
          // hi*hi products rotate over accumulators 0..2 (fewer sequential fp32 accumulations per accumulator:
          // the tensor core truncates when it adds), the small cross terms share accumulator 3
          const int am = unit % Cfg::NH;
          mma_tf32(tmem_d + am * C, ah, bh, idesc, (uint32_t)(unit >= Cfg::NH));   // first touch overwrites
          mma_tf32(tmem_d + 3 * C, ah, bl, idesc, (uint32_t)(unit > 0));
          mma_tf32(tmem_d + 3 * C, al, bh, idesc, 1);
        }
        mma_commit(&bar_empty[s]);   // frees the stage when these MMAs have read it
      }
      __syncwarp();
      if (c < 8) TC_MARK(8 + c);
    }
    if (elect_one()) mma_commit(&bar_done);
    __syncwarp();
    TC_MARK(3);
  };
  auto a_load = [&]() {
    // ---- activation tile: row p <-> (batch row, position - 7), transformed, split into tf32 hi / lo.
    //      One item = 4 channels of one row = one 16-byte core-matrix row: conflict-free vector stores,
    //      global loads coalesced along the row index; 4 items (16-32 loads) in flight per thread.
    constexpr int NIT = (C / 4) * ROWS, GI = 2 * ROWS;   // items, items per group
    constexpr int UN = 4;
    const int tin = J.tin;
    const bool affine = tin < 0 && J.x2 != nullptr;
    int garr = 0;
    for (int base = tid - LT0; base < NIT; base += UN * NLD) {
      float xv[UN][4], yv[UN][4];
      int kcs[UN];
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int idx = base + u * NLD;
        const int kc = idx / ROWS, p = idx - kc * ROWS;
        kcs[u] = kc;
        int bb, pos; bool ok;
        if (packed) { const int sgi = p / tl.seg; bb = b0 + sgi; pos = p - sgi * tl.seg - HALO; ok = sgi < tl.nseg && bb < P.B; }
        else { bb = b0; pos = l0 - HALO + p; ok = p < 128 + KW - 1; }
        const int xpos = J.xs * pos + J.xo;
        ok = ok && idx < NIT && pos >= 0 && xpos < J.xl;
        const long long XL = J.xl;
        const long long off = (long long)(4 * kc) * XL + xpos;
        const float* xr = J.x + (long long)bb * J.xbs + off;
        const float* yr = affine ? J.x2 + (long long)bb * J.x2bs + off : nullptr;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          xv[u][e] = ok ? xr[(long long)e * XL] : 0.f;
          yv[u][e] = (ok && affine) ? yr[(long long)e * XL] : 0.f;
        }
        if (!ok) kcs[u] = -1 - kc;
      }
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        const int idx = base + u * NLD;
        if (idx >= NIT) continue;
        const bool ok = kcs[u] >= 0;
        const int kc = ok ? kcs[u] : -1 - kcs[u];
        float hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ch = 4 * kc + e;
          float v = 0.f;
          if (ok) {
            if (tin == T_BN_RELU) { v = (xv[u][e] - sp[ch]) * sp[C + ch]; v = v > 0.f ? v : 0.f; }
            else if (tin == T_RELU) v = xv[u][e] > 0.f ? xv[u][e] : 0.f;
            else if (affine) v = fmaf(sp[ch], xv[u][e], fmaf(sp[C + ch], yv[u][e], sp[2 * C + ch]));
            else v = xv[u][e];
          }
          hi[e] = __uint_as_float(to_tf32(v));
          lo[e] = __uint_as_float(to_tf32(v - hi[e]));
        }
        *reinterpret_cast<float4*>(A_hi + (long long)idx * 4) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(A_lo + (long long)idx * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      }
      // groups whose items this thread has all written: publish (generic -> async proxy fence, then arrive)
      const int nxt = base + UN * NLD;
      while (garr < G && (garr + 1) * GI <= nxt) { fence_async_smem(); mbar_arrive(&bar_a[garr]); ++garr; }
    }
    while (garr < G) { fence_async_smem(); mbar_arrive(&bar_a[garr]); ++garr; }
  };
  constexpr int NPRE = NCH < NST ? NCH : NST;      // chunks that need no free-stage wait
  if (warp == 1) produce(0, NPRE);
  if (KW == 1 || warp >= 2) a_load();
  __syncwarp();
  if (warp == 1) produce(NPRE, NCH);
  else if (warp == 0) mma_loop();
  __syncwarp();

  // ---- epilogue: TMEM -> registers -> global (+ per-channel statistics)
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  TC_MARK(4);
  {
    const int row = 32 * (warp & 3) + lane;
    int bb, l; bool valid;
    if (packed) { const int sgi = row / tl.seg; bb = b0 + sgi; l = row - sgi * tl.seg; valid = sgi < tl.nseg && bb < P.B && l < L; }
    else { bb = b0; l = l0 + row; valid = l < L; }
    constexpr int CPW = C >= 64 ? C / 2 : 32;             // columns per warp (two warps share a TMEM lane quarter)
    const int col_base = C >= 64 ? (warp >> 2) * CPW : 0;
    const bool warp_on = C >= 64 || warp < 4;
    const int zpos = J.zs * l + J.zo;
    const long long ZL = J.zl;                                // channel pitch of z / mref
    valid = valid && zpos < J.zl;
    float* zrow = J.z + (long long)bb * J.zbs + zpos;
    const float* mrow = J.mref ? J.mref + (long long)bb * J.mrbs + zpos : nullptr;
    const int omode = J.omode;
    const bool stats = J.acc != nullptr;
    if (warp_on) {
#pragma unroll 1
      for (int c0 = col_base; c0 < col_base + CPW; c0 += 32) {
        float v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)c0;
        constexpr int NV = C < 32 ? C : 32;                 // valid channels of this 32-column block
        if constexpr (C >= 32) {
          tmem_ld32(taddr, v);
#pragma unroll
          for (int a = 1; a < kNAcc; ++a) {
            if (a >= Cfg::NH && a != kNAcc - 1) continue;       // accumulator never written (fewer than 3 units)
            float t[32];
            tmem_ld32(taddr + a * C, t);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += t[i];
          }
        } else {
          // C = 16: the four accumulators are the 16-column quarters of 64 allocated columns
          float t0[32], t1[32];
          tmem_ld32(taddr, t0);
          tmem_ld32(taddr + 32, t1);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = i < C ? (t0[i] + t0[C + i]) + ((Cfg::NH > 2 ? t1[i] : 0.f) + t1[C + i]) : 0.f;
        }
        float q[32];
        if (omode == O_STORE) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = valid ? v[i] : 0.f;
            if (valid && i < NV) zrow[(long long)(c0 + i) * ZL] = v[i];
            q[i] = v[i] * v[i];
          }
        } else if (omode == O_ACC) {          // later phases of a strided forward conv: add to the partial sum
          float old[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) old[i] = (valid && i < NV) ? zrow[(long long)(c0 + i) * ZL] : 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = valid ? old[i] + v[i] : 0.f;
            if (valid && i < NV) zrow[(long long)(c0 + i) * ZL] = v[i];
            q[i] = v[i] * v[i];
          }
        } else if (omode == O_RELU_ACC) {
          float xr[32], old[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            xr[i] = (valid && i < NV) ? mrow[(long long)(c0 + i) * ZL] : 0.f;
            old[i] = (valid && i < NV) ? zrow[(long long)(c0 + i) * ZL] : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (valid && i < NV && xr[i] > 0.f) zrow[(long long)(c0 + i) * ZL] = old[i] + v[i];
        } else {   // O_BNRELU_STORE
          float zi[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) zi[i] = (valid && i < NV) ? mrow[(long long)(c0 + i) * ZL] : 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[i] = (valid && i < NV && zi[i] > smean[i < NV ? c0 + i : 0]) ? v[i] : 0.f;
            if (valid && i < NV) zrow[(long long)(c0 + i) * ZL] = v[i];
            q[i] = v[i] * zi[i];
          }
        }
        if (stats && omode != O_RELU_ACC) {
          const float s = warp_transpose_sum(v, lane), qq = warp_transpose_sum(q, lane);
          if (lane < NV) {
            atomicAdd(&ssum[c0 + lane], (double)s);
            atomicAdd(&ssum[C + c0 + lane], (double)qq);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  TC_MARK(5);
  if (J.acc != nullptr)
    for (int i = tid; i < 2 * C; i += kTcThreads) atomicAdd(&J.acc[i], ssum[i]);
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TM_COLS) : "memory");
  }
  TC_MARK(6);
}

static inline bool tc_supported(int C) { return C == 16 || C == 32 || C == 64 || C == 128; }

template <int C, int KW>
static void launch_conv_tc_t(const TcBatch& b, dim3 grid, cudaStream_t st, const char* name) {
  const size_t smem = TcCfg<C, KW>::smem;
  static int prepared[kMaxDevices] = {0};
  if (need_prepare(prepared, (int)smem)) cudaFuncSetAttribute(k_conv_tc<C, KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  prof_before(name, "", st);
  k_conv_tc<C, KW><<<grid, dim3(kTcThreads), smem, st>>>(b);
  prof_after(st);
}

static int launch_conv_tc(TcBatch& b, cudaStream_t st) {
  if (b.n == 0) return 0;
  int tiles = 0;
  for (int i = 0; i < b.n; ++i) { const int t = TcTiling(b.j[i].L, b.KW).tiles(b.B); tiles = t > tiles ? t : tiles; }
  dim3 grid(tiles, 1, b.n);
  const int key = b.C * 100 + b.KW;
  switch (key) {
    case 1615: launch_conv_tc_t<16, 15>(b, grid, st, "k_conv_tc<16,15>"); break;
    case 1601: launch_conv_tc_t<16, 1>(b, grid, st, "k_conv_tc<16,1>"); break;
    case 3215: launch_conv_tc_t<32, 15>(b, grid, st, "k_conv_tc<32,15>"); break;
    case 6415: launch_conv_tc_t<64, 15>(b, grid, st, "k_conv_tc<64,15>"); break;
    case 12815: launch_conv_tc_t<128, 15>(b, grid, st, "k_conv_tc<128,15>"); break;
    case 3208: launch_conv_tc_t<32, 8>(b, grid, st, "k_conv_tc<32,8>"); break;
    case 6408: launch_conv_tc_t<64, 8>(b, grid, st, "k_conv_tc<64,8>"); break;
    case 12805: launch_conv_tc_t<128, 5>(b, grid, st, "k_conv_tc<128,5>"); break;
    case 3201: launch_conv_tc_t<32, 1>(b, grid, st, "k_conv_tc<32,1>"); break;
    case 6401: launch_conv_tc_t<64, 1>(b, grid, st, "k_conv_tc<64,1>"); break;
    case 12801: launch_conv_tc_t<128, 1>(b, grid, st, "k_conv_tc<128,1>"); break;
    default: set_error("conv_tc: unsupported (channels, taps)"); return GNAS_ERR_UNSUPPORTED;
  }
  b.n = 0;
  return 0;
}

}  // namespace tc
}  // namespace gnas
#endif  // !GNAS_EMUL

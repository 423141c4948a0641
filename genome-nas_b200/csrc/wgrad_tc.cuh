// Pointwise (1x1) conv weight gradients on the 5th-generation tensor cores (3xTF32, fp32 accumulate in TMEM):
//   dW[co][ci] += sum_{b,l} dz[b][co][l] * u[b][ci][l],   dz = A*g + Bz*z + D  (the BN backward folded into the load)
// for the pointwise convs of sep_conv / dil_conv (operations_14_9.py:36-64, backward).  Both operands are K-major as
// they lie in memory (positions contiguous).  128 / C jobs (edges x candidates of one cell node) are stacked along
// M and N so that one 128 x 128 x 8 MMA serves all of them; the epilogue keeps the diagonal C x C blocks.
// The kernel is loader-bound (three activation tensors are read once), not MMA-bound.
#pragma once
#include "tc_common.cuh"
#include "cnn_bwd.cuh"

#ifndef GNAS_EMUL
namespace gnas {
namespace tc {

// Per-job tap geometry of the conv weight gradient: the kernel's local tap k (0..15) belongs to weight tap
// tau0 + tstep * k (if k < ntaps), reads R at decimated position rs * pos + ro, and output position l pairs with
// input position l + k - halo.  Stride-1 conv_15: {7, 0, 1, 15, 1, 0}; phase phi of a stride-S conv: see cell.cu.
struct WgTap { int halo, tau0, tstep, ntaps, rs, ro; };
struct WgTcBatch { int n, B, rows_per_cta; WgJob j[kMaxDense]; WgTap tap[kMaxDense]; };

constexpr int kPwStages = 3;
constexpr int kPwOp = 8 * 128 * 4;                 // floats of one operand half: [8 k-chunks][128 rows][4]
constexpr int kPwStage = 4 * kPwOp;                // A_hi | A_lo | B_hi | B_lo
constexpr int kPwLoaders = kTcThreads - 32;        // warps 1..7
constexpr size_t kPwSmem = sizeof(float) * ((size_t)kPwStage * kPwStages + 3 * 128) + 128;

template <int C>
__global__ void __launch_bounds__(kTcThreads, 1) k_pw_wgrad_tc(const __grid_constant__ WgTcBatch P) {
  constexpr int NJ = 128 / C;                               // jobs stacked along M and N
  const int j0 = blockIdx.y * NJ;
  const int L = P.j[j0].Lo;
  const int spr = (L + 31) / 32;                            // stages per batch row
  const int bs = blockIdx.x * P.rows_per_cta, be = min(P.B, bs + P.rows_per_cta);
  if (bs >= be) return;
  const int nstage = (be - bs) * spr;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  extern __shared__ __align__(128) unsigned char pw_smem_raw[];
  float* sm = reinterpret_cast<float*>(pw_smem_raw);
  float* coef = sm + (size_t)kPwStage * kPwStages;          // [3][128]: A | Bz | D per stacked output channel
  __shared__ __align__(8) uint64_t bar_full[kPwStages], bar_empty[kPwStages], bar_done;
  __shared__ uint32_t tmem_base_s;
  constexpr uint32_t LBO = 128 * 16, SBO = 128;
  constexpr int TM_COLS = 512;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
#pragma unroll
    for (int i = 0; i < kPwStages; ++i) { mbar_init(&bar_full[i], kPwLoaders); mbar_init(&bar_empty[i], 1); }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 3 * 128; i += kTcThreads) {
    const int which = i >> 7, r = i & 127, jl = r / C, ch = r - jl * C;
    float v = which == 0 ? 1.f : 0.f;
    if (j0 + jl < P.n && P.j[j0 + jl].z != nullptr) v = P.j[j0 + jl].coef[which * C + ch];
    coef[i] = v;
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
      const int s = st % kPwStages;
      mbar_wait(&bar_full[s], (st / kPwStages) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t base = d0 + (uint64_t)(((uint32_t)s * kPwStage * 4) >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t koff = (uint64_t)((ks * 2 * LBO) >> 4);
          const uint64_t ah = base + koff, al = ah + (uint64_t)((kPwOp * 4) >> 4);
          const uint64_t bh = base + koff + (uint64_t)((2 * kPwOp * 4) >> 4), bl = bh + (uint64_t)((kPwOp * 4) >> 4);
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
    // ---- loaders: 4 consecutive positions of one channel -> (affine) -> tf32 hi / lo -> one 16-byte core-matrix row
    const int lt = tid - 32;
    constexpr int NIT = 2048, PER = (NIT + kPwLoaders - 1) / kPwLoaders;   // items per stage, per thread
    for (int st = 0; st < nstage; ++st) {
      const int s = st % kPwStages;
      if (st >= kPwStages) mbar_wait(&bar_empty[s], ((st / kPwStages) - 1) & 1);
      const int b = bs + st / spr, p0 = (st % spr) * 32;
      float* stage = sm + (size_t)s * kPwStage;
      float xv[PER][4], yv[PER][4];
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int idx = lt + i * kPwLoaders;
#pragma unroll
        for (int e = 0; e < 4; ++e) { xv[i][e] = 0.f; yv[i][e] = 0.f; }
        if (idx >= NIT) continue;
        const int op = idx >> 10, q = idx & 1023, rb = q >> 5, w = q & 31;
        const int r = (rb & 7) * 16 + (w & 15), jc = 2 * (rb >> 3) + (w >> 4);
        const int jl = r / C, ch = r - jl * C;
        if (j0 + jl >= P.n) continue;
        const WgJob& J = P.j[j0 + jl];
        const int pos = p0 + 4 * jc;
        if (op == 0) {
          const float* gp = J.g + (long long)b * J.gbs + (long long)ch * L + pos;
          const float* zp = J.z ? J.z + (long long)b * J.zbs + (long long)ch * L + pos : nullptr;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (pos + e < L) { xv[i][e] = gp[e]; if (zp) yv[i][e] = zp[e]; }
        } else {
          const float* rp = J.r + (long long)b * J.rbs + (long long)ch * L + pos;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (pos + e < L) xv[i][e] = rp[e];
        }
      }
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int idx = lt + i * kPwLoaders;
        if (idx >= NIT) continue;
        const int op = idx >> 10, q = idx & 1023, rb = q >> 5, w = q & 31;
        const int r = (rb & 7) * 16 + (w & 15), jc = 2 * (rb >> 3) + (w >> 4);
        const int pos = p0 + 4 * jc;
        float a[4], h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          a[e] = xv[i][e];
          if (op == 0) a[e] = (pos + e < L) ? fmaf(coef[r], xv[i][e], fmaf(coef[128 + r], yv[i][e], coef[256 + r])) : 0.f;
          h[e] = __uint_as_float(to_tf32(a[e]));
          l[e] = __uint_as_float(to_tf32(a[e] - h[e]));
        }
        float* dst = stage + (size_t)op * 2 * kPwOp + ((size_t)jc * 128 + r) * 4;
        *reinterpret_cast<float4*>(dst) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(dst + kPwOp) = make_float4(l[0], l[1], l[2], l[3]);
      }
      fence_async_smem();
      mbar_arrive(&bar_full[s]);
    }
  }
  __syncwarp();

  // ---- epilogue: TMEM -> registers -> red.add of the diagonal C x C blocks into the gradients
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  {
    const int q4 = warp & 3;
    const int row = 32 * q4 + lane;
    const int jl = row / C, co = row - jl * C;
    const bool have = j0 + jl < P.n;
    float* orow = have ? P.j[j0 + jl].dw + (size_t)co * C : nullptr;
    if constexpr (C >= 32) {
      constexpr int NB = C / 32;                              // 32-column blocks of the diagonal block (jl is warp-uniform)
      for (int blk = (warp >> 2); blk < NB; blk += 2) {
        const int c0 = jl * C + blk * 32;
        float v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(32 * q4) << 16) + (uint32_t)c0;
        tmem_ld32(taddr, v);
#pragma unroll
        for (int a = 1; a < 4; ++a) {
          float t[32];
          tmem_ld32(taddr + a * 128, t);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += t[i];
        }
        if (have) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) red_add_v4(orow + blk * 32 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
    } else if (warp < 4) {
      // several jobs per warp: the diagonal blocks of rows 32 q4 .. 32 q4 + 31 all lie in columns 32 q4 .. 32 q4 + 31
      float v[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(32 * q4) << 16) + (uint32_t)(32 * q4);
      tmem_ld32(taddr, v);
#pragma unroll
      for (int a = 1; a < 4; ++a) {
        float t[32];
        tmem_ld32(taddr + a * 128, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += t[i];
      }
      const int og_mine = (lane - lane % C) / C;
      float o[C];
#pragma unroll
      for (int og = 0; og < 32 / C; ++og)
        if (og == og_mine) {
#pragma unroll
          for (int i = 0; i < C; ++i) o[i] = v[og * C + i];
        }
      if (have) {
#pragma unroll
        for (int i = 0; i < C; i += 4) red_add_v4(orow + i, o[i], o[i + 1], o[i + 2], o[i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TM_COLS) : "memory");
  }
}

// Batch rows per CTA of a weight-gradient launch.  One CTA owns a whole SM (its operand stages fill the shared
// memory), so a launch costs ceil(CTAs / SMs) rounds of (rows * stages-per-row + fixed) stages: pick the row count
// that minimises it instead of a fixed chain length (e.g. 243 CTAs = one full round + one round at 64 %).
static inline int wgrad_rows_per_cta(int B, int ctas_per_row_split, int spr, int fixed_stages, int max_stages) {
  static int auto_on = -1;
  if (auto_on < 0) { const char* e = getenv("GNAS_WG_AUTO"); auto_on = e ? atoi(e) : 1; }
  if (!auto_on) return 0;
  int best = 1; long long best_cost = -1;
  for (int rows = 1; rows <= B; ++rows) {
    if (rows > 1 && rows * spr > max_stages) break;
    const long long grid = (long long)cdiv(B, rows) * ctas_per_row_split;
    const long long cost = cdivll(grid, 148) * ((long long)rows * spr + fixed_stages);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = rows; }
  }
  return best;
}

template <int C>
static void launch_pw_wgrad_tc_t(WgTcBatch& b, cudaStream_t st, const char* name) {
  static int prepared[kMaxDevices] = {0};
  if (need_prepare(prepared, (int)kPwSmem)) cudaFuncSetAttribute(k_pw_wgrad_tc<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPwSmem);
  const int groups = cdiv(b.n, 128 / C);
  int ksplit = cdiv(2 * 148, groups);
  ksplit = ksplit > b.B ? b.B : (ksplit < 1 ? 1 : ksplit);
  b.rows_per_cta = cdiv(b.B, ksplit);
  {
    const int spr = (b.j[0].Lo + 31) / 32;                   // 32-position stages per batch row
    const int r = wgrad_rows_per_cta(b.B, groups, spr, 4, 96);      // bounded chains: the tensor core truncates on accumulate
    if (r > 0) b.rows_per_cta = r;
  }
  prof_before(name, "", st);
  k_pw_wgrad_tc<C><<<dim3(cdiv(b.B, b.rows_per_cta), groups), dim3(kTcThreads), kPwSmem, st>>>(b);
  prof_after(st);
}

// jobs must share (Cout = Cin = C, stride 1, no input transform, Lo = Lin)
static inline bool wgrad_tc_channels(int C) { return C == 8 || C == 16 || C == 32 || C == 64 || C == 128; }
static inline bool pw_wgrad_tc_supported(const WgBatch& b) {
  if (!wgrad_tc_channels(b.Cout)) return false;
  for (int i = 0; i < b.n; ++i) {
    const WgJob& j = b.j[i];
    if (j.Cin != b.Cout || j.tin != T_NONE || j.Lo != j.Lin || j.Lo != b.j[0].Lo) return false;
  }
  return true;
}

static int launch_pw_wgrad_tc(WgBatch& b, cudaStream_t st) {
  if (b.n == 0) return 0;
  WgTcBatch t; t.n = b.n; t.B = b.B;
  for (int i = 0; i < b.n; ++i) t.j[i] = b.j[i];
  switch (b.Cout) {
    case 8: launch_pw_wgrad_tc_t<8>(t, st, "k_pw_wgrad_tc<8>"); break;
    case 16: launch_pw_wgrad_tc_t<16>(t, st, "k_pw_wgrad_tc<16>"); break;
    case 32: launch_pw_wgrad_tc_t<32>(t, st, "k_pw_wgrad_tc<32>"); break;
    case 64: launch_pw_wgrad_tc_t<64>(t, st, "k_pw_wgrad_tc<64>"); break;
    default: launch_pw_wgrad_tc_t<128>(t, st, "k_pw_wgrad_tc<128>"); break;
  }
  b.n = 0;
  return 0;
}

// ------------------------------------------------------------------ conv_15 weight gradient (stride 1)
//   dW[co][ci][t] += sum_{b,l} dz[b][co][l] * R[b][ci][l + t - 7],    R = ReLU(x) or ReLU(BN(z1))
// Positions are the reduction (K) dimension, so a tap is a shift along K; K-major operands only shift in whole
// 16-byte chunks (4 positions).  The other three shifts are put into the rows: the A tile holds RA = 128 / C copies
// of dz shifted by sa = 0..RA-1 positions, the B tile RB = 4 / RA copies of R shifted by RA * sb, and the MMA for
// tap group a reads B four positions (one chunk) further on:  t = sa + RA * sb + 4 * a  (t = 15 is discarded).
// C = 128 has RB = 4 row blocks of 128: one CTA per block (blockIdx.z).  C = 16 / 8 carry 8 / 16 sub-shifts in the A
// rows, so a tap group covers 8 / 16 taps and the groups are 2 / 4 chunks apart.
template <int C> struct CwCfg {
  static constexpr int RA = 128 / C;                        // sub-shifts carried by the A rows
  static constexpr int RB = RA >= 4 ? 1 : 4 / RA;           // sub-shifts carried by the B rows
  static constexpr int TS = RA * RB;                        // taps per tap group (4, 8 or 16)
  static constexpr int NG = 16 / TS;                        // tap groups = accumulators
  static constexpr int CO = TS / 4;                         // chunk offset between tap groups
  static constexpr int NB = C >= 64 ? 128 : (C >= 16 ? C : 16);   // B rows per CTA (= MMA N; C = 8 pads to 16)
  static constexpr int NZ = C == 128 ? 4 : 1;               // CTAs along the sb dimension
  static constexpr int BCH = 8 + (NG - 1) * CO;             // B chunks per stage
  static constexpr int A_OP = 8 * 128 * 4, B_OP = BCH * NB * 4;
  static constexpr int STAGE = 2 * A_OP + 2 * B_OP;         // floats
  static constexpr int NST = C >= 64 ? 2 : 3;
  static constexpr int NV = NG * NB * 4 <= 512 ? 4 : 1;     // accumulator variants: hi*hi rotates over 3, cross terms get 1
  static constexpr int TM_COLS = NV * NG * NB < 32 ? 32 : NV * NG * NB;
  static constexpr size_t smem = sizeof(float) * ((size_t)STAGE * NST + 3 * 128 + 2 * 128) + 128;
};

template <int C>
__global__ void __launch_bounds__(kTcThreads, 1) k_conv_wgrad_tc(const __grid_constant__ WgTcBatch P) {
  using Cfg = CwCfg<C>;
  constexpr int RA = Cfg::RA, RB = Cfg::RB, TS = Cfg::TS, NG = Cfg::NG, CO = Cfg::CO, NB = Cfg::NB, BCH = Cfg::BCH, NST = Cfg::NST;
  constexpr int A_OP = Cfg::A_OP, B_OP = Cfg::B_OP, STAGE = Cfg::STAGE, NV = Cfg::NV;
  const WgJob& J = P.j[blockIdx.y];
  const WgTap TP = P.tap[blockIdx.y];
  const int L = J.Lo, LR = J.Lin;                           // output positions (K), input row length
  const int nguse = (TP.ntaps + TS - 1) / TS < NG ? (TP.ntaps + TS - 1) / TS : NG;   // tap groups that hold wanted taps
  const int spr = (L + RA - 1 + 31) / 32;                   // stages per batch row (A rows are shifted by up to RA - 1)
  const int bs = blockIdx.x * P.rows_per_cta, be = min(P.B, bs + P.rows_per_cta);
  if (bs >= be) return;
  const int nstage = (be - bs) * spr;
  const int sb0 = C == 128 ? blockIdx.z : 0;                // first B row block of this CTA
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  extern __shared__ __align__(128) unsigned char cw_smem_raw[];
  float* sm = reinterpret_cast<float*>(cw_smem_raw);
  float* coef = sm + (size_t)STAGE * NST;                   // [3][128]: A | Bz | D per A row
  float* bpar = coef + 3 * 128;                             // [2][128]: mean | rstd per B row
  __shared__ __align__(8) uint64_t bar_full[NST], bar_empty[NST], bar_done;
  __shared__ uint32_t tmem_base_s;
  constexpr uint32_t LBO_A = 128 * 16, LBO_B = NB * 16, SBO = 128;
  constexpr int TM_COLS = Cfg::TM_COLS;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
#pragma unroll
    for (int i = 0; i < NST; ++i) { mbar_init(&bar_full[i], kPwLoaders); mbar_init(&bar_empty[i], 1); }
    mbar_init(&bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 3 * 128; i += kTcThreads) {
    const int which = i >> 7, ch = (i & 127) % C;
    coef[i] = J.z != nullptr ? J.coef[which * C + ch] : (which == 0 ? 1.f : 0.f);
  }
  for (int i = tid; i < 2 * 128; i += kTcThreads) {
    const int which = i >> 7, ch = (i & 127) % C;
    bpar[i] = J.tin == T_BN_RELU ? J.r_stat[which * C + ch] : (which == 0 ? 0.f : 1.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  if (warp == 0) {
    // ---- MMA issuer
    const uint32_t idesc = make_idesc(NB);
    const uint64_t a0 = make_desc(smem_u32(sm), LBO_A, SBO), b0 = make_desc(smem_u32(sm + 2 * A_OP), LBO_B, SBO);
    for (int st = 0; st < nstage; ++st) {
      const int s = st % NST;
      mbar_wait(&bar_full[s], (st / NST) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t so = (uint64_t)(((uint32_t)s * STAGE * 4) >> 4);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t ah = a0 + so + (uint64_t)((ks * 2 * LBO_A) >> 4), al = ah + (uint64_t)((A_OP * 4) >> 4);
#pragma unroll
          for (int a = 0; a < NG; ++a) {
            if (a >= nguse) continue;
            const uint64_t bh = b0 + so + (uint64_t)(((ks * 2 + a * CO) * LBO_B) >> 4), bl = bh + (uint64_t)((B_OP * 4) >> 4);
            if constexpr (NV == 1) {
              const uint32_t acc = tmem_d + a * NB;
              mma_tf32(acc, ah, bh, idesc, (uint32_t)(st | ks));     // first touch overwrites
              mma_tf32(acc, ah, bl, idesc, 1);
              mma_tf32(acc, al, bh, idesc, 1);
            } else {
              // room in TMEM: hi*hi rotates over three accumulators, the cross terms share a fourth (shorter
              // truncating accumulation chains, as in the forward kernels)
              const int unit = st * 4 + ks;
              mma_tf32(tmem_d + ((unit % 3) * NG + a) * NB, ah, bh, idesc, (uint32_t)(unit >= 3));
              mma_tf32(tmem_d + (3 * NG + a) * NB, ah, bl, idesc, (uint32_t)(unit > 0));
              mma_tf32(tmem_d + (3 * NG + a) * NB, al, bh, idesc, 1);
            }
          }
        }
        mma_commit(&bar_empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) mma_commit(&bar_done);
    __syncwarp();
  } else {
    // ---- loaders
    const int lt = tid - 32;
    constexpr int NA = 8 * 128, NBI = BCH * NB, NIT = NA + NBI, PER = (NIT + kPwLoaders - 1) / kPwLoaders;
    const int tin = J.tin;
    for (int st = 0; st < nstage; ++st) {
      const int s = st % NST;
      if (st >= NST) mbar_wait(&bar_empty[s], ((st / NST) - 1) & 1);
      const int b = bs + st / spr, p0 = (st % spr) * 32;
      float* stage = sm + (size_t)s * STAGE;
      const float* gb = J.g + (long long)b * J.gbs;
      const float* zb = J.z ? J.z + (long long)b * J.zbs : nullptr;
      const float* rb_ = J.r + (long long)b * J.rbs;
      float xv[PER][4], yv[PER][4];
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int idx = lt + i * kPwLoaders;
#pragma unroll
        for (int e = 0; e < 4; ++e) { xv[i][e] = 0.f; yv[i][e] = 0.f; }
        if (idx >= NIT) continue;
        if (idx < NA) {                       // A: row (sa, co), chunk jc: dz[co][p0 + 4 jc + e - sa]
          const int rbk = idx >> 5, w = idx & 31;
          const int r = (rbk & 7) * 16 + (w & 15), jc = 2 * (rbk >> 3) + (w >> 4);
          const int sa = r / C, ch = r - sa * C;
          const int pos = p0 + 4 * jc - sa;
          const float* gp = gb + (long long)ch * L + pos;
          const float* zp = zb ? zb + (long long)ch * L + pos : nullptr;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (pos + e >= 0 && pos + e < L) { xv[i][e] = gp[e]; if (zp) yv[i][e] = zp[e]; }
        } else {                              // B: row (sb, ci), chunk jc: R[ci][p0 + 4 jc + e + RA sb - 7]
          const int q = idx - NA;
          const int jc = q / NB, r = q - jc * NB;
          const int sbl = r / C, ch = r - sbl * C;
          const int pos = p0 + 4 * jc + RA * (sb0 + sbl) - TP.halo;
          const float* rp = rb_ + (long long)ch * LR + TP.ro;
          if (sbl < RB || C == 128) {             // C = 8: rows 8..15 are padding
#pragma unroll
            for (int e = 0; e < 4; ++e)
              if (pos + e >= 0 && TP.rs * (pos + e) + TP.ro < LR) xv[i][e] = rp[(long long)TP.rs * (pos + e)];
          }
        }
      }
#pragma unroll
      for (int i = 0; i < PER; ++i) {
        const int idx = lt + i * kPwLoaders;
        if (idx >= NIT) continue;
        float a[4], h[4], l[4];
        float* dst;
        if (idx < NA) {
          const int rbk = idx >> 5, w = idx & 31;
          const int r = (rbk & 7) * 16 + (w & 15), jc = 2 * (rbk >> 3) + (w >> 4);
          const int pos = p0 + 4 * jc - r / C;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            a[e] = (pos + e >= 0 && pos + e < L) ? fmaf(coef[r], xv[i][e], fmaf(coef[128 + r], yv[i][e], coef[256 + r])) : 0.f;
          dst = stage + ((size_t)jc * 128 + r) * 4;
        } else {
          const int q = idx - NA;
          const int jc = q / NB, r = q - jc * NB;
          const int pos = p0 + 4 * jc + RA * (sb0 + r / C) - TP.halo;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float v = xv[i][e];
            if (tin == T_BN_RELU) v = (v - bpar[r]) * bpar[128 + r];
            if (tin == T_RELU || tin == T_BN_RELU) v = v > 0.f ? v : 0.f;
            a[e] = (pos + e >= 0 && TP.rs * (pos + e) + TP.ro < LR && (r / C < RB || C == 128)) ? v : 0.f;
          }
          dst = stage + 2 * A_OP + ((size_t)jc * NB + r) * 4;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) { h[e] = __uint_as_float(to_tf32(a[e])); l[e] = __uint_as_float(to_tf32(a[e] - h[e])); }
        *reinterpret_cast<float4*>(dst) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(dst + (idx < NA ? A_OP : B_OP)) = make_float4(l[0], l[1], l[2], l[3]);
      }
      fence_async_smem();
      mbar_arrive(&bar_full[s]);
    }
  }
  __syncwarp();

  // ---- epilogue: D_a[(sa, co)][(sb, ci)] -> dW[co][ci][sa + RA sb + TS a]
  mbar_wait(&bar_done, 0);
  tc_fence_after();
  {
    const int q4 = warp & 3;
    const int row = 32 * q4 + lane;
    const int sa = row / C, co = row - sa * C;
    float* wrow = J.dw + (size_t)co * C * 15;
    if constexpr (NB >= 32) {
      constexpr int NBLK = NB / 32;
      for (int u = (warp >> 2); u < nguse * NBLK; u += 2) {
        const int a = u / NBLK, blk = u - a * NBLK;
        float v[32];
        const uint32_t taddr = tmem_d + ((uint32_t)(32 * q4) << 16) + (uint32_t)(a * NB + blk * 32);
        tmem_ld32(taddr, v);
#pragma unroll
        for (int vv = 1; vv < NV; ++vv) {
          float t[32];
          tmem_ld32(taddr + vv * NG * NB, t);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += t[i];
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = blk * 32 + i;
          const int sbl = col / C, ci = col - sbl * C;
          const int t = sa + RA * (sb0 + sbl) + TS * a, tau = TP.tau0 + TP.tstep * t;
          if (t < TP.ntaps && tau < 15) atomicAdd(wrow + (size_t)ci * 15 + tau, v[i]);
        }
      }
    } else if (warp < 4) {
      constexpr int NC = NG * NB;                             // accumulator columns per variant (16 or 32)
      constexpr int NLD = (NV * NC + 31) / 32;
      float v[NLD][32];
#pragma unroll
      for (int k = 0; k < NLD; ++k) tmem_ld32(tmem_d + ((uint32_t)(32 * q4) << 16) + (uint32_t)(32 * k), v[k]);
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        float sum = 0.f;
#pragma unroll
        for (int vv = 0; vv < NV; ++vv) sum += v[(vv * NC + i) / 32][(vv * NC + i) % 32];
        const int a = i / NB, ci = i - a * NB;
        const int t = sa + TS * a, tau = TP.tau0 + TP.tstep * t;
        if (ci < C && t < TP.ntaps && tau < 15) atomicAdd(wrow + (size_t)ci * 15 + tau, sum);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TM_COLS) : "memory");
  }
}

template <int C>
static void launch_conv_wgrad_tc_t(WgTcBatch& b, cudaStream_t st, const char* name) {
  using Cfg = CwCfg<C>;
  static int prepared[kMaxDevices] = {0};
  if (need_prepare(prepared, (int)Cfg::smem)) cudaFuncSetAttribute(k_conv_wgrad_tc<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem);
  const int L = b.j[0].Lo, spr = (L + Cfg::RA - 1 + 31) / 32;
  static int chain = -1;                     // stages per accumulation chain (GNAS_WG_CHAIN: A/B runs)
  if (chain < 0) { const char* e = getenv("GNAS_WG_CHAIN"); chain = e ? atoi(e) : 10; }
  int rows = chain / spr;                    // ~10 stages (= 40 k-steps) per accumulation chain
  rows = rows < 1 ? 1 : rows;
  if (!getenv("GNAS_WG_CHAIN")) {
    const int r = wgrad_rows_per_cta(b.B, b.n * Cfg::NZ, spr, 4, 40);      // chains beyond ~40 stages lose to the truncation bias / parallelism
    if (r > 0) rows = r;
  }
  b.rows_per_cta = rows;
  prof_before(name, "", st);
  k_conv_wgrad_tc<C><<<dim3(cdiv(b.B, rows), b.n, Cfg::NZ), dim3(kTcThreads), Cfg::smem, st>>>(b);
  prof_after(st);
}

// jobs must share (Cout = Cin = C, Lo); stride 1: Lo = Lin
static inline bool conv_wgrad_tc_supported(const WgBatch& b, int stride) {
  if (!wgrad_tc_channels(b.Cout)) return false;
  if (stride > 1 && b.n * stride > kMaxDense) return false;
  for (int i = 0; i < b.n; ++i) {
    const WgJob& j = b.j[i];
    if (j.Cin != b.Cout || (stride == 1 && j.Lo != j.Lin) || j.Lo != b.j[0].Lo) return false;
    if (!(j.tin == T_RELU || j.tin == T_BN_RELU || j.tin == T_NONE)) return false;
  }
  return true;
}

// stride S > 1: one job per input phase phi = 0..S-1 (x[S q + phi]): weight taps tau = (phi + 7) mod S + S k
static int launch_conv_wgrad_tc(WgBatch& b, int stride, cudaStream_t st) {
  if (b.n == 0) return 0;
  WgTcBatch t; t.n = 0; t.B = b.B;
  for (int i = 0; i < b.n; ++i)
    for (int phi = 0; phi < stride; ++phi) {
      t.j[t.n] = b.j[i];
      WgTap& tp = t.tap[t.n++];
      if (stride == 1) { tp.halo = 7; tp.tau0 = 0; tp.tstep = 1; tp.ntaps = 15; tp.rs = 1; tp.ro = 0; }
      else {
        const int t0 = (phi + 7) % stride;
        tp.halo = (7 + phi - t0) / stride; tp.tau0 = t0; tp.tstep = stride; tp.ntaps = (14 - t0) / stride + 1;
        tp.rs = stride; tp.ro = phi;
      }
    }
  switch (b.Cout) {
    case 8: launch_conv_wgrad_tc_t<8>(t, st, "k_conv_wgrad_tc<8>"); break;
    case 16: launch_conv_wgrad_tc_t<16>(t, st, "k_conv_wgrad_tc<16>"); break;
    case 32: launch_conv_wgrad_tc_t<32>(t, st, "k_conv_wgrad_tc<32>"); break;
    case 64: launch_conv_wgrad_tc_t<64>(t, st, "k_conv_wgrad_tc<64>"); break;
    default: launch_conv_wgrad_tc_t<128>(t, st, "k_conv_wgrad_tc<128>"); break;
  }
  b.n = 0;
  return 0;
}

}  // namespace tc
}  // namespace gnas
#endif  // !GNAS_EMUL

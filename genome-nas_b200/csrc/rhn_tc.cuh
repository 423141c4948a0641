// Recurrent highway layer, forward, on the 5th-generation tensor cores (error-compensated 3xTF32).
// Replaces the recurrent products of darts_tools/model_discCNN.py:255-268 ([x_t*xm | h*hm] W0) and
// model_search.py:68-79 ((s_j*hm) Ws[i] for every edge (i, j)) -- the per-step batch x nhid GEMMs north_star (2)
// asks to put on tensor-core MMA -- inside ONE persistent cooperative kernel.
//
// Decomposition.  CTA g owns NC = 4 hidden columns n0 = 4g .. 4g+3, i.e. the 8 output columns (c | h) of W0 and
// of each Ws[i].  A time step is 9 dependent stages: stage 0 produces s_0 from [x_t | h_{t-1}], stage s = j + 1
// multiplies the new state s_j with the slices of Ws[j..7] ("multiply on arrival": the edge (j, j) completes
// node j -> s_{j+1}, the other 7 - j products are parked in the accumulators of their nodes).  Per stage:
//     D[64 batch rows][R columns] = S[64][K = H] * Wslice[K][R],   R = 8 (8 - j)  (8 for stage 0, K = 2H)
// as tcgen05.mma.kind::tf32 with M = 64 (batch), N = R, K = 8 per instruction; both operands are pre-split into
// tf32 hi + lo and three products (hi*hi, hi*lo, lo*hi) accumulate in fp32 in tensor memory.  The tensor core
// TRUNCATES when it adds into the accumulator; in a 378-stage recurrence that bias does not average out (one
// 64-step chain per stage: output error 2.1e-5 at T = 42 against 5.9e-6 for fp32 FFMA, measured and reproduced
// by a truncation model).  So the chains are kept to ONE 64-column K chunk (8 products in each of three
// accumulators: hi*hi, hi*lo, lo*hi), and the epilogue warps drain every chunk into fp32 registers (round-to-
// nearest adds) while the next chunk multiplies into the second accumulator set: fp32-class error in the model.
// Measured on a B200 (tools/umma_probe.cu): an M = 64 instruction costs 27 cycles for N <= 32 and 33 for N = 64
// (the shared-memory read of A), so the batch is the M side.
//
// Data movement.  Nothing but the operands of the next MMAs is ever staged: the masked state is PUBLISHED by its
// producer already split into hi | lo and laid out as the K-major no-swizzle UMMA operand ([4-column group][batch]
// [4]), so a consumer fetches a 64-column K chunk (32 KB) with one TMA bulk copy into a 3-slot ring; the weight
// slices are packed once per call (k_rhn_tc_pack) block-major ([chunk][matrix i = 7..0][hi|lo][4-column group][8
// rows][4]: SBO = one block, LBO = 128 bytes), so the 8 (8 - j) rows stage j needs are one contiguous range.
// BatchNorm over the batch is per column, hence CTA-local; a grid barrier separates stages.
//
// Warp roles (448 threads): warps 0-7 epilogue (TMEM -> registers -> shared, gates, BatchNorm, publish, stash for
// backward), warp 8 state producer (grid-barrier wait + TMA), warps 9 and 10 MMA issuers (one elected lane each;
// they alternate accumulator groups and own one accumulator set each, so the barrier polls and bookkeeping of one
// overlap the issue of the other -- a single issuing thread is latency-bound), warps 11-13 weight loaders (cp.async: a 1-D bulk copy costs ~600 cycles of the TMA unit whatever its size -- measured, two
// per chunk made the unit the bottleneck -- so only the state goes through it).
// The same source is compiled for the CPU emulator against a functional model of mbarrier / bulk copy / tcgen05
// (tests/emul/tc_emul.h), which is how the layouts and hand-offs are tested without a GPU.
#pragma once
#include "tc_common.cuh"

namespace gnas {
namespace rtc {

using namespace gnas::tc;

constexpr int NC = 4;                    // hidden columns per CTA
constexpr int BPT = 64;                  // batch rows of the MMA (padded)
constexpr int KCH = 64;                  // K per pipeline chunk
constexpr int KC4 = KCH / 4;             // 4-column groups (core-matrix columns) per chunk
constexpr int NSLOT = 3;
constexpr int GRP = 1;                   // chunks multiplied into one accumulator set before it is drained (power of two)
constexpr int S_HALF = KC4 * BPT * 4;    // floats of the hi (or lo) half of a state chunk (16 KB)
constexpr int S_CHUNK = 2 * S_HALF;      // 32 KB
constexpr int R0 = 2 * NC;               // output columns of one weight matrix per CTA (c | h)
constexpr int RMAX = 8 * R0;             // stage 1: all eight Ws
constexpr int W_BLOCK = 2 * KC4 * R0 * 4;  // floats of one matrix block of a chunk: [hi|lo][KC4][R0][4] (4 KB)
constexpr int W_CHUNK_MAX = 8 * W_BLOCK;
constexpr int SLOT = S_CHUNK + W_CHUNK_MAX;          // floats per ring slot (64 KB)
constexpr int PITCH = 72;                // part[column][batch] pitch: conflict-free for the dump and for (batch, column) readers
constexpr int EPI = 256;                 // epilogue threads: (batch row m = tid / 4, column = tid % 4)
constexpr int NLOAD = NSLOT;             // weight-loader warps: loader w owns ring slot w (chunks w, w + 3, ...), so it can
                                         // never run two mbarrier phases ahead of the consumer (parity aliasing)
constexpr int THREADS = EPI + 32 * (3 + NLOAD);   // + state producer, two MMA issuers, the weight loaders
constexpr int NSET = 2;                  // accumulator sets in tensor memory: {hi*hi, hi*lo, lo*hi} x 64 columns, pitch 256
constexpr int TM_COLS = NSET * 256;
static_assert(NSET == 2, "the set / parity arithmetic below assumes two accumulator sets");

struct Plan {
  int nch;                               // K chunks per H
  long long xc, ringc, wp, total;        // extra forward workspace (floats): x operands, state ring, packed weights
  long long wp_cta;                      // packed floats per CTA
};
static inline Plan plan(int T, int H) {
  Plan p;
  p.nch = H / KCH;
  long long o = 0;
  p.xc = o; o += (long long)T * p.nch * S_CHUNK;
  p.ringc = o; o += 10LL * p.nch * S_CHUNK;
  p.wp_cta = (long long)p.nch * (2 * W_BLOCK + 8 * W_BLOCK);
  p.wp = o; o += (long long)(H / NC) * p.wp_cta;
  p.total = o;
  return p;
}
static inline bool supported(int B, int H, int nsm) { return B >= 1 && B <= BPT && H % KCH == 0 && H / NC <= nsm; }

// offset (floats) of element (hidden column n, batch row m) inside a [nch][hi|lo][KC4][BPT][4] operand buffer
__device__ __forceinline__ long long op_off(int n, int m, int lo) {
  const int g4 = n >> 2;
  return (long long)(g4 / KC4) * S_CHUNK + (long long)lo * S_HALF + ((g4 % KC4) * BPT + m) * 4 + (n & 3);
}

// ------------------------------------------------------------------ per-call preparation
// x operands of every time step and the initial hidden state, masked, split and laid out (from the feature-major
// buffers the prologue of rhn.cu wrote)
struct Pre { int T, H, BP; const float* xmT; const float* h0T; const float* hmT; float* xc; float* ring9; };
static __global__ void __launch_bounds__(256) k_rhn_tc_pre(const __grid_constant__ Pre P) {
  const long long HB = (long long)P.H * BPT;
  const long long total = (long long)(P.T + 1) * HB;
  const long long slab_f = (long long)(P.H / KCH) * S_CHUNK;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / HB);
    const int r = (int)(i - (long long)t * HB);
    const int m = r % BPT, n = r / BPT;
    float v = 0.f;
    if (m < P.BP) v = t < P.T ? P.xmT[((long long)t * P.H + n) * P.BP + m] : P.h0T[(long long)n * P.BP + m] * P.hmT[(long long)n * P.BP + m];
    const float hi = __uint_as_float(to_tf32(v));
    float* dst = t < P.T ? P.xc + (long long)t * slab_f : P.ring9;
    dst[op_off(n, m, 0)] = hi;
    dst[op_off(n, m, 1)] = __uint_as_float(to_tf32(v - hi));
  }
}

// weights -> per-CTA operand slices, block-major: [W0 x part: nch x block] [W0 h part: nch x block]
// [Ws: nch x 8 blocks, block b = matrix i = 7 - b], block = [hi|lo][KC4][R0 rows = (c|h, column)][4].
// Stage j reads the first 8 - j blocks of a chunk.
struct Pack { int H; const float* W0; const float* Ws; float* wp; long long wp_cta; };
static __global__ void __launch_bounds__(256) k_rhn_tc_pack(const __grid_constant__ Pack P) {
  const int H = P.H, nch = H / KCH;
  const int g = blockIdx.x, n0 = g * NC;
  float* out = P.wp + (long long)g * P.wp_cta;
  const int half = KC4 * R0 * 4;                    // one of hi / lo of a block
  const int nblk = 2 * nch + 8 * nch;               // blocks of this CTA in packing order
  for (int e = threadIdx.x; e < nblk * half; e += blockDim.x) {
    const int blk = e / half, q = e - blk * half;
    const int c4 = q & 3, rr = (q >> 2) % R0, kcc = (q >> 2) / R0;
    const int col = (rr / NC) * H + n0 + (rr % NC);
    float v;
    if (blk < 2 * nch) {
      const int k = blk * KCH + kcc * 4 + c4;         // blk = part * nch + chunk: rows 0..2H-1 of W0 in order
      v = P.W0[(long long)k * 2 * H + col];
    } else {
      const int chunk = (blk - 2 * nch) / 8, i = 7 - (blk - 2 * nch) % 8;
      const int k = chunk * KCH + kcc * 4 + c4;
      v = P.Ws[((long long)i * H + k) * 2 * H + col];
    }
    const float hi = __uint_as_float(to_tf32(v));
    float* dst = out + (long long)blk * W_BLOCK + q;
    dst[0] = hi;
    dst[half] = __uint_as_float(to_tf32(v - hi));
  }
}

// ------------------------------------------------------------------ the persistent forward kernel
struct Fwd {
  int T, B, BP, H, training;
  float momentum;
  const float* hmT; const float* h0T; const float* probs;
  const float* xc; float* ringc; const float* wp; long long wp_cta;
  float* ST; float* CHT; float* C0T; float* HT; float* stat;
  float* out; float* running;
  unsigned* bar;
  RhnTable tb;
};

struct Job { int t, stage, chunk, first, last, gfirst, glast; };     // chunk: index inside the stage; g*: accumulator group
// Jobs (chunks) in issue order: per time step, stage 0 = 2 nch chunks (x then h), stages 1..8 = nch chunks each.
// Every role walks the same sequence with this cursor: no divisions in the single-thread loops (a 64-bit
// division per chunk cost the MMA warp ~400 cycles), ring slot / phase kept incrementally.
struct Cursor {
  int nch, slot, round;        // ring slot of the job and how often the ring has wrapped
  int g;                       // accumulator group of the job (GRP consecutive chunks of one stage)
  long long q;
  Job j;
  __device__ __forceinline__ explicit Cursor(int nch_) : nch(nch_), slot(0), round(0), g(0), q(0) {
    j.t = 0; j.stage = 0; j.chunk = 0; j.first = 1; j.last = 2 * nch_ == 1;
    j.gfirst = 1; j.glast = GRP == 1 || j.last;
  }
  __device__ __forceinline__ void next() {
    if (j.glast) ++g;
    const int n = j.stage == 0 ? 2 * nch : nch;
    if (++j.chunk == n) { j.chunk = 0; if (++j.stage == 9) { j.stage = 0; ++j.t; } }
    j.first = j.chunk == 0;
    j.last = j.chunk == (j.stage == 0 ? 2 * nch : nch) - 1;
    j.gfirst = (j.chunk & (GRP - 1)) == 0;
    j.glast = (j.chunk & (GRP - 1)) == GRP - 1 || j.last;
    if (++slot == NSLOT) { slot = 0; ++round; }
    ++q;
  }
  __device__ __forceinline__ uint32_t full_parity() const { return (uint32_t)(round & 1); }          // this use of the slot
  __device__ __forceinline__ uint32_t prev_parity() const { return (uint32_t)((round - 1) & 1); }    // the use before it
};
__device__ __forceinline__ int rows_of(int stage) { return stage == 0 ? R0 : R0 * (9 - stage); }

__device__ __forceinline__ void spin_until(const unsigned* bar, unsigned target) {
  const long long t0 = gnas_clock();
  while (ld_acquire_u32(bar) < target) {
#ifdef GNAS_EMUL
    GNAS_SPIN_PAUSE();
#else
    if (gnas_clock() - t0 > 4000000000LL) __trap();      // ~2 s: never leave the GPU spinning
#endif
  }
  (void)t0;
}

#ifdef GNAS_RHN_TIMING
// clock64 marks of CTA 0 at time step 1 (tools/rhn_timing.py): [stage][0 first chunk complete, 1 drained + dumped,
// 2 published (barrier arrive), 3 stage end, 4 producer passed the grid barrier, 5 MMA warp saw the first chunk,
// 6 MMA warp issued the last chunk, 7 BatchNorm + stores issued]
#define RTC_MARK(stage, k) do { if (blockIdx.x == 0 && t_mark == 1) dbg[(stage) * 8 + (k)] = clock64(); } while (0)
#else
#define RTC_MARK(stage, k) do { } while (0)
#endif

__global__ void __launch_bounds__(THREADS, 1) k_rhn_fwd_tc(const __grid_constant__ Fwd P) {
#ifdef GNAS_RHN_TIMING
  long long* dbg = reinterpret_cast<long long*>(P.bar + 16);
#endif
  const int H = P.H, B = P.B, T = P.T, nch = H / KCH;
  const int n0 = blockIdx.x * NC;
  const int tid = threadIdx.x, lane = tid & 31;
#ifdef GNAS_EMUL
  const int warp = tid >> 5;
#else
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
#endif
  const long long HB = (long long)H * P.BP;            // stash tensors keep the caller's padded batch pitch
  GNAS_DYN_SMEM(float, smem);
  float* ring = smem;                                    // NSLOT x { S chunk | W chunk }
  float* part = ring + (size_t)NSLOT * SLOT;             // [RMAX][PITCH] D of the stage, column-major
  float* pp = part + RMAX * PITCH;                       // [36][4] probabilities
  float* pq = pp + 36 * 4;                               // [36]
  float* red = pq + 36;                                  // [2][8 warps][NC] BatchNorm partials
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 2 * 8 * NC + 4);   // 8-byte aligned: all counts above are even
  uint64_t* bar_full = bars;                             // [NSLOT] state bytes + weight bytes of a chunk
  uint64_t* bar_empty = bars + NSLOT;                    // [NSLOT] MMAs of the slot retired
  uint64_t* bar_acc_full = bars + 2 * NSLOT;             // [NSET] accumulator set of a chunk complete
  uint64_t* bar_acc_empty = bars + 2 * NSLOT + NSET;     // [NSET] ... and drained into registers by all 8 epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSLOT + 2 * NSET);
  const long long njobs = (long long)T * 10 * nch;

  if (tid < 36) {
    float q = 0.f;
    for (int k = 0; k < 4; ++k) {
      const int ix = P.tb.pidx[tid][k];
      const float v = ix >= 0 ? P.probs[ix] : 0.f;
      pp[tid * 4 + k] = v; q += v;
    }
    pq[tid] = q;
  }
  if (warp == 0) tmem_alloc(tmem_slot, TM_COLS);
  if (tid == 32) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&bar_full[i], 2); mbar_init(&bar_empty[i], 1); }
    for (int i = 0; i < NSET; ++i) { mbar_init(&bar_acc_full[i], 1); mbar_init(&bar_acc_empty[i], 8); }
    mbar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const float* wp = P.wp + (long long)blockIdx.x * P.wp_cta;

  if (warp == 8) {
    // ================= state producer: one lane; TMA bulk copy of a 32 KB operand chunk per job
    if (lane == 0) {
      for (Cursor cu(nch); cu.q < njobs; cu.next()) {
        const int s = cu.slot;
        const Job j = cu.j;
        if (cu.round > 0) mbar_wait(&bar_empty[s], cu.prev_parity());
        const float* src;
        if (j.stage == 0 && j.chunk < nch) src = P.xc + ((long long)j.t * nch + j.chunk) * S_CHUNK;
        else {
          // barrier number = publishes completed: s_{stage-1} of this step is publish t*9 + stage; h_{t-1} is t*9
          const int slot = j.stage == 0 ? 9 : j.stage - 1;
          const int c = j.stage == 0 ? j.chunk - nch : j.chunk;
          const long long nb = (long long)j.t * 9 + j.stage;
          if (nb > 0 && c == 0) {
            spin_until(P.bar, (unsigned)(nb * gridDim.x));
            fence_proxy_async_all();
#ifdef GNAS_RHN_TIMING
            { const int t_mark = j.t; RTC_MARK(j.stage, 4); }
#endif
          }
          src = P.ringc + ((long long)slot * nch + c) * S_CHUNK;
        }
        mbar_expect_tx(&bar_full[s], S_CHUNK * 4);
        bulk_g2s(ring + (size_t)s * SLOT, src, S_CHUNK * 4, &bar_full[s]);
      }
    }
  } else if (warp >= 11) {
    // ================= weight loaders (NLOAD warps, chunk q goes to loader q % NLOAD): the 8 - j leading blocks of
    // the chunk are one contiguous range; cp.async 16 bytes per lane, then wait for the own group, fence for the
    // async proxy (the tensor core reads shared memory through it) and hand the slot over.  A loader never waits
    // for anything but its slot, so the hand-over of a chunk does not depend on later chunks; one chunk takes a
    // warp ~2 k cycles (issue + L2 round trip), hence one loader per ring slot.
    Cursor cu(nch);
    for (int i = 0; i < warp - 11 && cu.q < njobs; ++i) cu.next();
    while (cu.q < njobs) {
      const int s = cu.slot;
      const Job j = cu.j;
      if (cu.round > 0 && lane == 0) mbar_wait(&bar_empty[s], cu.prev_parity());
      __syncwarp();
      const float* src = (j.stage == 0 ? wp + (long long)j.chunk * W_BLOCK
                                       : wp + 2LL * nch * W_BLOCK + (long long)j.chunk * 8 * W_BLOCK) + lane * 4;
      const int nblk = j.stage == 0 ? 1 : 9 - j.stage;
      float* dst = ring + (size_t)s * SLOT + S_CHUNK + lane * 4;
      for (int b = 0; b < nblk; ++b) {
#pragma unroll
        for (int i = 0; i < W_BLOCK / 128; ++i) cp_async16(dst + b * W_BLOCK + i * 128, src + b * W_BLOCK + i * 128);
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[s]);
      for (int i = 0; i < NLOAD && cu.q < njobs; ++i) cu.next();
    }
  } else if (warp == 9 || warp == 10) {
    // ================= MMA issuers: warp 9 takes the even groups (accumulator set 0), warp 10 the odd ones (set 1);
    // every chunk starts fresh accumulators
    const int mine = warp == 9 ? 0 : 1;
    for (Cursor cu(nch); cu.q < njobs; cu.next()) {
      if ((cu.g & 1) != mine) continue;
      const int s = cu.slot, set = cu.g & 1;
      const Job j = cu.j;
      const int R = rows_of(j.stage);
      if (lane == 0) {            // one lane polls (32 pollers on one mbarrier only slow the arrivals down)
        if (j.gfirst && cu.g >= NSET) mbar_wait(&bar_acc_empty[set], (uint32_t)(((cu.g >> 1) - 1) & 1));
#ifdef GNAS_RHN_TIMING
        if (blockIdx.x == 0 && j.t == 1 && j.stage == 1) dbg[72 + j.chunk * 5 + 0] = clock64();
#endif
        mbar_wait(&bar_full[s], cu.full_parity());
#ifdef GNAS_RHN_TIMING
        if (blockIdx.x == 0 && j.t == 1 && j.stage == 1) dbg[72 + j.chunk * 5 + 1] = clock64();
#endif
      }
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
#ifdef GNAS_RHN_TIMING
        { const int t_mark = j.t; if (j.first || (j.stage == 0 && j.chunk == nch)) RTC_MARK(j.stage, 5); }
#endif
        const uint32_t idesc = make_idesc_mn(64, R);
        const uint32_t sbase = smem_u32(ring + (size_t)s * SLOT);
        const uint32_t wbase = sbase + S_CHUNK * 4;
        const uint32_t acc = tmem_d + (uint32_t)set * 256;
        constexpr uint32_t lbo_a = BPT * 16, lbo_b = R0 * 16, sbo_b = W_BLOCK * 4, w_lo = KC4 * R0 * 16;
        // descriptors differ only in the start-address field (bits 0..13, units of 16 bytes): build the four bases
        // once per chunk and step them with plain adds -- the single issuing thread is latency-bound, every
        // instruction between two MMAs shows (measured: 53 cycles per MMA with make_desc inside the loop)
        const uint64_t a_hi = make_desc(sbase, lbo_a, 128), a_lo = a_hi + (uint64_t)((S_HALF * 4) >> 4);
        const uint64_t b_hi = make_desc(wbase, lbo_b, sbo_b), b_lo = b_hi + (uint64_t)(w_lo >> 4);
#pragma unroll
        for (int ks = 0; ks < KCH / 8; ++ks) {
          const uint64_t ka = (uint64_t)(ks * ((2 * lbo_a) >> 4)), kb = (uint64_t)(ks * ((2 * lbo_b) >> 4));
          // three accumulators in turn: back-to-back MMAs into the SAME accumulator would serialise on its latency
          const uint32_t accf = (uint32_t)(ks > 0 || !j.gfirst);     // fresh accumulators at the first K step of a group
          mma_tf32(acc, a_hi + ka, b_hi + kb, idesc, accf);
          mma_tf32(acc + 64, a_hi + ka, b_lo + kb, idesc, accf);
          mma_tf32(acc + 128, a_lo + ka, b_hi + kb, idesc, accf);
        }
        mma_commit(&bar_empty[s]);
        if (j.glast) mma_commit(&bar_acc_full[set]);
#ifdef GNAS_RHN_TIMING
        { const int t_mark = j.t; if (j.last) RTC_MARK(j.stage, 6); }
        if (blockIdx.x == 0 && j.t == 1 && j.stage == 1) dbg[72 + j.chunk * 5 + 2] = clock64();
#endif
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    // ================= epilogue: thread = (batch row m, own column col)
    const int m = tid >> 2, col = tid & 3, n = n0 + col;
    const bool real = m < B;
    const int q4 = warp & 3, hf = warp >> 2;            // TMEM lane quarter / column half of this warp
    const float hmask = real ? P.hmT[(long long)n * P.BP + m] : 0.f;
    float hprev = real ? P.h0T[(long long)n * P.BP + m] : 0.f;
    float run_mean = 0.f, run_var = 1.f;
    if (P.running) { run_mean = P.running[n]; run_var = P.running[H + n]; }
    float accn[8], sl[9];
#pragma unroll
    for (int i = 0; i < 8; ++i) accn[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) sl[i] = 0.f;
    const float invB = 1.f / (float)B;
    int flip = 0;
    // sum over the batch of `v` for this thread's column (fixed order: deterministic); all 256 threads take part
    auto batch_sum = [&](float v) -> float {
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      float* r = red + flip * 8 * NC;
      flip ^= 1;
      if (lane < NC) r[warp * NC + lane] = v;
      named_bar_sync(1, EPI);
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += r[w * NC + col];
      return t;
    };
    // BatchNorm over the batch of the raw state -> s (own copy), stats stash, publish masked hi | lo into ring slot
    auto bn_publish = [&](float raw, int t, int stage, int slot) -> float {
      float mean, rstd;
      if (P.training) {
        mean = batch_sum(real ? raw : 0.f) * invB;
        const float d = real ? raw - mean : 0.f;
        const float var = batch_sum(d * d) * invB;
        rstd = 1.0f / sqrtf(var + kBnEps);
        run_mean = (1.f - P.momentum) * run_mean + P.momentum * mean;
        run_var = (1.f - P.momentum) * run_var + P.momentum * (B > 1 ? var * (float)B / (float)(B - 1) : var);
      } else {
        mean = run_mean;
        rstd = 1.0f / sqrtf(run_var + kBnEps);
      }
      const float v = real ? (raw - mean) * rstd : 0.f;
      if (m == 0) { float* st = P.stat + ((long long)t * 9 + stage) * 2 * H; st[n] = mean; st[H + n] = rstd; }
      if (m < P.BP) P.ST[((long long)t * 9 + stage) * HB + (long long)n * P.BP + m] = v;
      if (slot >= 0) {
        const float mv = v * hmask;
        const float hi = __uint_as_float(to_tf32(mv));
        float* dst = P.ringc + (long long)slot * nch * S_CHUNK;
        dst[op_off(n, m, 0)] = hi;
        dst[op_off(n, m, 1)] = __uint_as_float(to_tf32(mv - hi));
      }
      return v;
    };
    auto publish_done = [&]() {
      named_bar_sync(2, EPI);
      if (tid == 0) { __threadfence(); atomicAdd(P.bar, 1u); }
    };
#ifdef GNAS_RHN_TIMING
#define RTC_PUBLISHED(stage) do { const int t_mark = t; if (tid == 0) RTC_MARK(stage, 2); } while (0)
#else
#define RTC_PUBLISHED(stage) do { } while (0)
#endif
    // contribution of edge (i, j) with pre-activations c | h to node i
    auto edge = [&](int i, int j, float c, float h) -> float {
      const int p = i * (i + 1) / 2 + j;
      const float q = pq[p];
      const float p0 = pp[p * 4], p1 = pp[p * 4 + 1], p2 = pp[p * 4 + 2], p3 = pp[p * 4 + 3];
      if (p0 == 0.f && p1 == 0.f && p2 == 0.f && p3 == 0.f) return 0.f;
      const float sj = sl[j];
      const float g = sigmoidf_(c);
      float F = 0.f;
      if (p0 != 0.f) F = fmaf(p0, tanhf(h), F);
      if (p1 != 0.f) F = fmaf(p1, h, F);
      if (p2 != 0.f) F = fmaf(p2, sigmoidf_(h), F);
      if (p3 != 0.f) F = fmaf(p3, h > 0.f ? h : 0.f, F);
      return q * sj + g * (F - q * sj);
    };

    long long ge = 0;                                   // accumulator-group counter, same order as the MMA warps
    for (int t = 0; t < T; ++t) {
      for (int stage = 0; stage < 9; ++stage) {
        const int R = rows_of(stage);
        // ---- drain every chunk's accumulators into fp32 registers (round-to-nearest), then -> shared memory
        const int njs = stage == 0 ? 2 * nch : nch;
        float racc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) racc[i] = 0.f;
        for (int c = 0; c < njs; c += GRP, ++ge) {
          const int set = (int)(ge & 1);
          if (lane == 0) mbar_wait(&bar_acc_full[set], (uint32_t)((ge >> 1) & 1));
          __syncwarp();
          tc_fence_after();
#ifdef GNAS_RHN_TIMING
          { const int t_mark = t; if (tid == 0 && c == (stage == 0 ? nch : 0)) RTC_MARK(stage, 0); }
          if (blockIdx.x == 0 && t == 1 && stage == 1 && tid == 0) dbg[72 + c * 5 + 3] = clock64();
#endif
          if (32 * hf < R) {
            const uint32_t taddr = tmem_d + ((uint32_t)(32 * q4) << 16) + (uint32_t)(set * 256 + 32 * hf);
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              float v[32];
              tmem_ld32(taddr + a * 64, v);
#pragma unroll
              for (int i = 0; i < 32; ++i) racc[i] += v[i];
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_acc_empty[set]);
#ifdef GNAS_RHN_TIMING
          if (blockIdx.x == 0 && t == 1 && stage == 1 && tid == 0) dbg[72 + c * 5 + 4] = clock64();
#endif
        }
        if (32 * hf < R && lane < 16) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (32 * hf + i < R) part[(32 * hf + i) * PITCH + 16 * q4 + lane] = racc[i];
        }
        named_bar_sync(3, EPI);
#ifdef GNAS_RHN_TIMING
        const int t_mark = t;
        if (tid == 0) RTC_MARK(stage, 1);
#endif
        if (stage == 0) {
          // ---- s_0 = BN(h + sig(c0) (tanh(h0) - h))     (model_discCNN.py:258-268)
          const float c = part[col * PITCH + m], h = part[(NC + col) * PITCH + m];
          if (m < P.BP) {
            P.C0T[((long long)t * 2 + 0) * HB + (long long)n * P.BP + m] = real ? c : 0.f;
            P.C0T[((long long)t * 2 + 1) * HB + (long long)n * P.BP + m] = real ? h : 0.f;
          }
          const float raw = hprev + sigmoidf_(c) * (tanhf(h) - hprev);
          sl[0] = bn_publish(raw, t, 0, 0);
#ifdef GNAS_RHN_TIMING
          if (tid == 0) RTC_MARK(0, 7);
#endif
          publish_done();
          RTC_PUBLISHED(0);
        } else {
          const int j = stage - 1;
          // ---- critical: edge (j, j) completes node j -> s_{j+1}   (model_search.py:68-93)
          const int rb = (7 - j) * R0;                      // its rows are the last block of the operand
          const float cc = part[(rb + col) * PITCH + m], hh = part[(rb + NC + col) * PITCH + m];
          float raw = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i == j) { accn[i] += edge(i, j, cc, hh); raw = accn[i]; accn[i] = 0.f; }
          const float v = bn_publish(raw, t, stage, stage < 8 ? stage : -1);
#ifdef GNAS_RHN_TIMING
          if (tid == 0) RTC_MARK(stage, 7);
#endif
#pragma unroll
          for (int i = 1; i < 9; ++i) if (i == stage) sl[i] = v;
          if (stage == 8) {
            // h_t = mean(s_1..s_8)   (model_search.py:96); its masked copy feeds stage 0 of the next step
            float a = 0.f;
#pragma unroll
            for (int i = 1; i <= 8; ++i) a += sl[i];
            a = real ? a * 0.125f : 0.f;
            hprev = a;
            const float mv = a * hmask;
            const float hi = __uint_as_float(to_tf32(mv));
            float* dst = P.ringc + 9LL * nch * S_CHUNK;
            dst[op_off(n, m, 0)] = hi;
            dst[op_off(n, m, 1)] = __uint_as_float(to_tf32(mv - hi));
            if (m < P.BP) P.HT[(long long)t * HB + (long long)n * P.BP + m] = a;
            if (real) P.out[((long long)t * B + m) * H + n] = a;
          }
          publish_done();
          RTC_PUBLISHED(stage);
          // ---- off the critical path: stash the pre-activations, park the other products of s_j in their nodes
          if (m < P.BP) {
            float* chp = P.CHT + (((long long)t * 36 + j * (j + 1) / 2 + j) * 2) * HB + (long long)n * P.BP + m;
            chp[0] = real ? cc : 0.f; chp[HB] = real ? hh : 0.f;
          }
#pragma unroll
          for (int b = 0; b < 7; ++b) {
            const int i = 7 - b;
            if (i > j) {
              const float c = part[(b * R0 + col) * PITCH + m], h = part[(b * R0 + NC + col) * PITCH + m];
              if (m < P.BP) {
                float* chp = P.CHT + (((long long)t * 36 + i * (i + 1) / 2 + j) * 2) * HB + (long long)n * P.BP + m;
                chp[0] = real ? c : 0.f; chp[HB] = real ? h : 0.f;
              }
              accn[i] += edge(i, j, c, h);
            }
          }
        }
        // `part` is rewritten by the next stage's dump: everyone must be done reading it
        named_bar_sync(4, EPI);
#ifdef GNAS_RHN_TIMING
        if (tid == 0) RTC_MARK(stage, 3);
#endif
      }
    }
    if (P.training && P.running && m == 0) { P.running[n] = run_mean; P.running[H + n] = run_var; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_d, TM_COLS); }
}

static inline size_t smem_bytes() {
  return sizeof(float) * ((size_t)NSLOT * SLOT + RMAX * PITCH + 36 * 4 + 36 + 2 * 8 * NC + 4) + 8 * (2 * NSLOT + 2 * NSET) + 16;
}

}  // namespace rtc
}  // namespace gnas

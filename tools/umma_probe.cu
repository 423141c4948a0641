// Planning probe for the tensor-core RHN recurrence (not part of the product):
//   1. issue cost of tcgen05.mma kind::tf32 (SS mode, no-swizzle K-major operands) for M in {64,128}, N in {8..256}
//   2. which TMEM lane holds which row of D for M = 64
//   3. TMA bulk-copy rate L2 -> shared memory per SM when G CTAs stream the same 256 KB (the state broadcast)
//   4. cost of a software grid barrier (atomicAdd + ld.acquire spin) with G co-resident CTAs
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe tools/umma_probe.cu
#include "../genome-nas_b200/csrc/common.cuh"
namespace gnas {
void prof_before(const char*, const char*, cudaStream_t) {}
void prof_after(cudaStream_t) {}
void set_error(const char* m) { fprintf(stderr, "error: %s\n", m); }
int check_launch(const char*) { return 0; }
}
#include "../genome-nas_b200/csrc/tc_common.cuh"
#include <vector>
using namespace gnas;
using namespace gnas::tc;

__device__ __forceinline__ uint32_t idesc_mn(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- 1. MMA issue cost
__global__ void __launch_bounds__(128, 1) k_rate(long long* out, int M, int N, int nmma, int distinct) {
  extern __shared__ __align__(128) unsigned char raw[];
  float* sm = reinterpret_cast<float*>(raw);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024; i += 128) sm[i] = 0.001f * (float)(i % 97);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tmem_s;
  long long t0 = 0, t1 = 0;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t id = idesc_mn(M, N);
      // A: [2 k-chunks][128 rows][4] at 0; B: [2][256 rows][4] at 64 KB
      const uint64_t a0 = make_desc(smem_u32(sm), 128 * 16, 128);
      const uint64_t b0 = make_desc(smem_u32(sm) + 64 * 1024, 256 * 16, 128);
      t0 = clock64();
      for (int i = 0; i < nmma; ++i) {
        // distinct: walk over 8 different K steps of the staged operands (what a real K loop does)
        const uint64_t off = distinct ? (uint64_t)(((i & 7) * 2 * 128 * 16) >> 4) : 0;
        const uint64_t offb = distinct ? (uint64_t)(((i & 7) * 2 * 256 * 16) >> 4) : 0;
        mma_tf32(td + (i & 1) * 256, a0 + off, b0 + offb, id, i >= 2);
      }
      mma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (elect_one()) { out[blockIdx.x * 2] = t0; out[blockIdx.x * 2 + 1] = t1; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(td), "r"(512) : "memory"); }
}

// ---- 2. M = 64 layout: D[r][n] = r + 1
__global__ void __launch_bounds__(128, 1) k_layout(float* out, int M) {
  extern __shared__ __align__(128) unsigned char raw[];
  float* sm = reinterpret_cast<float*>(raw);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* A = sm;            // [2 kc][128 rows][4]
  float* Bm = sm + 2 * 128 * 4;   // [2 kc][8 rows][4]
  for (int i = tid; i < 2 * 128 * 4 + 2 * 8 * 4; i += 128) sm[i] = 0.f;
  __syncthreads();
  if (tid < M) A[tid * 4] = (float)(tid + 1);      // k = 0 of row tid
  if (tid < 8) Bm[tid * 4] = 1.f;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(32) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t td = tmem_s;
  // clear the accumulator region first (all 128 lanes x 32 columns) so that untouched lanes read 0
  {
    const uint32_t taddr = td + ((uint32_t)(32 * warp) << 16);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    if (elect_one()) {
      mma_tf32(td, make_desc(smem_u32(A), 128 * 16, 128), make_desc(smem_u32(Bm), 8 * 16, 128), idesc_mn(M, 8), 1);
      mma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t r[8];
  const uint32_t taddr = td + ((uint32_t)(32 * warp) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 8; ++c) out[(32 * warp + lane) * 8 + c] = __uint_as_float(r[c]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(td), "r"(32) : "memory"); }
}

// ---- 3. bulk copies L2 -> smem: every CTA streams the same `total` bytes in `chunk`-byte pieces, `depth` in flight
__global__ void __launch_bounds__(128, 1) k_bulk(const float* src, long long* out, int total, int chunk, int depth, int reps, int offset_per_cta) {
  extern __shared__ __align__(128) unsigned char raw[];
  __shared__ __align__(8) uint64_t bar[8];
  const int tid = threadIdx.x;
  if (tid == 0) { for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const char* s = reinterpret_cast<const char*>(src) + (size_t)blockIdx.x * offset_per_cta;
    const int nch = total / chunk;
    t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      const int base = rep * nch;
      for (int c = 0; c < nch + depth; ++c) {
        if (c >= depth) { const int w = base + c - depth; mbar_wait(&bar[w % depth], (w / depth) & 1); }
        if (c < nch) {
          const int w = base + c;
          mbar_expect_tx(&bar[w % depth], chunk);
          bulk_g2s(raw + (size_t)(w % depth) * chunk, s + (size_t)c * chunk, chunk, &bar[w % depth]);
        }
      }
    }
    t1 = clock64();
    out[blockIdx.x * 2] = t0; out[blockIdx.x * 2 + 1] = t1;
  }
}

// ---- 4. software grid barrier
__global__ void __launch_bounds__(128, 1) k_gbar(unsigned* bar, long long* out, int n) {
  long long t0 = clock64();
  unsigned target = 0;
  for (int i = 0; i < n; ++i) {
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(bar, 1u);
      long long w0 = clock64();
      for (;;) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        if (v >= target) break;
        if (clock64() - w0 > 2000000000LL) break;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = t0; out[blockIdx.x * 2 + 1] = clock64(); }
}

int main() {
  long long* d; cudaMalloc(&d, 4096 * 8);
  std::vector<long long> h(4096);
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("== 1. tcgen05.mma kind::tf32 SS, cycles per MMA (512 back to back, 148 CTAs; same operands | 8 distinct K steps)\n");
  for (int M : {64, 128})
    for (int N : {8, 16, 32, 64, 128, 256})
      for (int distinct : {0, 1}) {
        const int n = 512;
        k_rate<<<148, 128, 200 * 1024>>>(d, M, N, n, distinct);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("M=%d N=%d: %s\n", M, N, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h.data(), d, 148 * 16, cudaMemcpyDeviceToHost);
        double s = 0; for (int i = 0; i < 148; ++i) s += (double)(h[2 * i + 1] - h[2 * i]);
        printf("M=%3d N=%3d %s: %.1f cycles/MMA\n", M, N, distinct ? "distinct" : "same    ", s / 148 / n);
      }
  printf("== 2. TMEM lane of each D row, M = 64 and 128 (value = row + 1 in column 0)\n");
  float* df; cudaMalloc(&df, 128 * 8 * 4);
  std::vector<float> hf(128 * 8);
  for (int M : {64, 128}) {
    k_layout<<<1, 128, 8192>>>(df, M);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("layout M=%d: %s\n", M, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(hf.data(), df, 128 * 8 * 4, cudaMemcpyDeviceToHost);
    printf("M=%d: lane->row: ", M);
    for (int l = 0; l < 128; ++l) printf("%d%s", (int)hf[l * 8] - 1, l % 32 == 31 ? " | " : ",");
    printf("\n   columns of lane 0: "); for (int c = 0; c < 8; ++c) printf("%g ", hf[c]); printf("\n");
  }
  printf("== 3. TMA bulk L2->smem: cycles to stream 256 KB per CTA (same source for all CTAs), after one warm-up pass\n");
  float* src; cudaMalloc(&src, 64 << 20); cudaMemset(src, 0, 64 << 20);
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int G : {1, 64, 128, 148})
    for (int chunk : {8192, 16384, 32768})
      for (int depth : {2, 4})
        for (int offs : {0, 262144}) {
          if (chunk * depth > 160 * 1024) continue;
          k_bulk<<<G, 128, 200 * 1024>>>(src, d, 262144, chunk, depth, 8, offs);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("bulk: %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(h.data(), d, G * 16, cudaMemcpyDeviceToHost);
          double s = 0, mx = 0; for (int i = 0; i < G; ++i) { double v = (double)(h[2 * i + 1] - h[2 * i]) / 8; s += v; mx = v > mx ? v : mx; }
          printf("G=%3d chunk=%5d depth=%d %s: mean %.0f max %.0f cycles per 256 KB (%.1f B/clk/SM)\n", G, chunk, depth,
                 offs ? "distinct src" : "shared src  ", s / G, mx, 262144.0 / (s / G));
        }
  printf("== 4. software grid barrier, cycles per barrier (1000 barriers)\n");
  unsigned* bar; cudaMalloc(&bar, 64);
  for (int G : {32, 64, 128, 148}) {
    cudaMemset(bar, 0, 64);
    void* args[] = {(void*)&bar, (void*)&d, nullptr};
    int n = 1000; args[2] = (void*)&n;
    cudaError_t e = cudaLaunchCooperativeKernel((void*)k_gbar, dim3(G), dim3(128), args, 0, 0);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("gbar: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d, G * 16, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < G; ++i) s += (double)(h[2 * i + 1] - h[2 * i]);
    printf("G=%3d: %.0f cycles per barrier\n", G, s / G / n);
  }
  return 0;
}

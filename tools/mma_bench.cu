// Micro-benchmark: issue rate of legacy mma.sync.m16n8k8 TF32 vs FFMA on this GPU (planning data for the
// 3xTF32 path; not part of the product).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_mma(float* out, int iters) {
  unsigned a0 = threadIdx.x, a1 = 2, a2 = 3, a3 = 4, b0 = 5, b1 = 6;
  float c[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_ffma(float* out, int iters) {
  float c[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) c[i] = (float)i;
  float a = 1.0001f + threadIdx.x * 1e-9f, b = 0.5f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i] = fmaf(c[i], a, b);
  }
  float s = 0.f;
  for (int i = 0; i < 32; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int iters = 20000, blocks = 148 * 2;
    k_mma<<<blocks, warps * 32>>>(out, 100);
    cudaEventRecord(e0);
    k_mma<<<blocks, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)blocks * warps * iters * 8 * (2.0 * 16 * 8 * 8);
    printf("mma.sync tf32 m16n8k8: %d warps/CTA x2 CTA/SM: %.1f TFLOP/s\n", warps, flops / ms / 1e9);
    k_ffma<<<blocks, warps * 32>>>(out, 100);
    cudaEventRecord(e0);
    k_ffma<<<blocks, warps * 32>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    flops = (double)blocks * warps * 32 * iters * 32 * 2.0;
    printf("ffma fp32:             %d warps/CTA x2 CTA/SM: %.1f TFLOP/s\n", warps, flops / ms / 1e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

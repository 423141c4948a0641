// Stand-alone timing of the tcgen05 conv_15 kernel with in-kernel phase marks (tuning aid, not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -DGNAS_TC_MARKS -o tools/tc_bench tools/tc_bench.cu
#include "../genome-nas_b200/csrc/common.cuh"
namespace gnas {
void prof_before(const char*, const char*, cudaStream_t) {}
void prof_after(cudaStream_t) {}
void set_error(const char* m) { fprintf(stderr, "error: %s\n", m); }
int check_launch(const char*) { return 0; }
}
#include "../genome-nas_b200/csrc/conv15_tc.cuh"
#include <vector>
#include <cstdlib>
using namespace gnas;

int main(int argc, char** argv) {
  const int C = argc > 1 ? atoi(argv[1]) : 64, L = argc > 2 ? atoi(argv[2]) : 125, B = argc > 3 ? atoi(argv[3]) : 64;
  const int jobs = argc > 4 ? atoi(argv[4]) : 1, mode = argc > 5 ? atoi(argv[5]) : 0, KW = argc > 6 ? atoi(argv[6]) : 15;   // mode 0 fwd, 1 dx (BNRELU_STORE)
  const size_t n = (size_t)B * C * L;
  float *x, *x2, *z, *w, *stat, *mref; double* acc;
  cudaMalloc(&x, n * 4 * jobs); cudaMalloc(&x2, n * 4 * jobs); cudaMalloc(&z, n * 4 * jobs); cudaMalloc(&mref, n * 4 * jobs);
  cudaMalloc(&w, (size_t)2 * C * C * 15 * 4); cudaMalloc(&stat, 3 * C * 4); cudaMalloc(&acc, 2 * C * 8);
  std::vector<float> h(n * jobs);
  for (auto& v : h) v = (float)rand() / RAND_MAX - 0.5f;
  cudaMemcpy(x, h.data(), n * 4 * jobs, cudaMemcpyHostToDevice); cudaMemcpy(x2, h.data(), n * 4 * jobs, cudaMemcpyHostToDevice);
  cudaMemcpy(mref, h.data(), n * 4 * jobs, cudaMemcpyHostToDevice);
  cudaMemcpy(w, h.data(), (size_t)2 * C * C * 15 * 4 < n * 4 * jobs ? (size_t)2 * C * C * 15 * 4 : n * 4 * jobs, cudaMemcpyHostToDevice);
  cudaMemset(stat, 0, 3 * C * 4); cudaMemset(acc, 0, 2 * C * 8);
  tc::TcBatch b; b.B = B; b.C = C; b.KW = KW;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 20;
  float ms = 0;
  for (int r = 0; r < reps + 3; ++r) {
    if (r == 3) cudaEventRecord(e0);
    b.n = jobs;
    for (int i = 0; i < jobs; ++i) {
      tc::TcJob& t = b.j[i];
      t.x = x + i * n; t.xbs = (long long)C * L; t.x2 = mode ? x2 + i * n : nullptr; t.x2bs = (long long)C * L; t.in_stat = stat; t.w = w;
      t.z = z + i * n; t.zbs = (long long)C * L; t.acc = acc; t.mref = mode ? mref + i * n : nullptr; t.mrbs = (long long)C * L; t.mstat = stat;
      t.L = L; t.tin = mode ? -1 : (KW == 1 ? T_NONE : T_BN_RELU); t.omode = mode ? (KW == 1 ? O_STORE : O_BNRELU_STORE) : O_STORE;
      if (mode && KW == 1) { t.acc = nullptr; t.mref = nullptr; }
    }
    tc::launch_conv_tc(b, 0);
  }
  cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t err = cudaGetLastError();
  printf("C=%d L=%d B=%d jobs=%d mode=%d KW=%d: %.1f us/launch (%s) tiles=%d\n", C, L, B, jobs, mode, KW, 1e3 * ms / reps, cudaGetErrorString(err),
         tc::TcTiling(L, KW).tiles(B) * jobs);
#ifdef GNAS_TC_MARKS
  long long m[64];
  cudaMemcpyFromSymbol(m, tc::g_tc_marks, sizeof(m));
  printf("  cycles: alloc+init %lld | A tile %lld | main loop %lld | drain %lld | epilogue %lld | stats+dealloc %lld | total %lld\n",
         m[1] - m[0], m[2] - m[1], m[3] - m[2], m[4] - m[3], m[5] - m[4], m[6] - m[5], m[6] - m[0]);
  printf("  first chunk iterations:");
  for (int i = 0; i < 8; ++i) printf(" %lld", m[8 + i] - (i ? m[8 + i - 1] : m[2]));
  printf("\n");
#endif
  return 0;
}

// Development probe: per-phase clock counters of the small-n Jacobi kernel (compiles csrc/eigh_jacobi.cu with
// GDFT_EIG_PROFILE) + total kernel time, for a random symmetric n x n matrix.
#define GDFT_EIG_PROFILE 1
#include "../graddft_b200/csrc/eigh_jacobi.cu"
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace gdft { thread_local int g_last_cuda_error = 0; std::atomic<unsigned long long> g_launches{0}; }
int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 43, batch = 2;
  std::vector<double> h(batch * n * n);
  srand(n);
  for (int b = 0; b < batch; b++)
    for (int i = 0; i < n; i++)
      for (int j = 0; j <= i; j++) { double v = rand() / (double)RAND_MAX - 0.5; h[b * n * n + i * n + j] = h[b * n * n + j * n + i] = v; }
  double *A, *w, *V; cudaMalloc(&A, h.size() * 8); cudaMalloc(&w, batch * n * 8); cudaMalloc(&V, h.size() * 8);
  cudaMemcpy(A, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  for (int i = 0; i < 3; i++) gdft_sym_eigh(0, batch, n, A, w, V);
  cudaDeviceSynchronize();
  long long zero[8] = {0}; cudaMemcpyToSymbol(gdft::g_eig_prof, zero, sizeof(zero));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); gdft_sym_eigh(0, batch, n, A, w, V); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long p[8]; cudaMemcpyFromSymbol(p, gdft::g_eig_prof, sizeof(p));
  printf("n=%d: kernel %.1f us, rounds %lld; per round clocks: rotation %.0f, barrier0 %.0f, A update %.0f, barrier1 %.0f | V update %.0f, V idle before barrier0 %.0f\n",
         n, ms * 1e3, p[4], (double)p[0] / p[4], (double)p[1] / p[4], (double)p[2] / p[4], (double)p[3] / p[4], (double)p[5] / p[4], (double)p[6] / p[4]);
  return 0;
}

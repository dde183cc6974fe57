// Development probe: dependent-chain latency (SM clocks per op) of FP64 DFMA / rsqrt / sqrt / divide and of a
// __syncthreads round trip on B200; sizes the serial part of a Jacobi round (csrc/eigh_jacobi.cu).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double* out, long long* clk, double x0, int iters) {
  double x = x0 + threadIdx.x * 1e-12;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (OP == 0) x = fma(x, 0.999999, 1e-7);
    if (OP == 1) x = rsqrt(x + 1.5);
    if (OP == 2) x = sqrt(x + 1.5);
    if (OP == 3) x = 1.7 / (x + 1.5);
    if (OP == 4) { __syncthreads(); x += 1.0; }
    if (OP == 5) x = (double)rsqrtf((float)(x + 1.5));
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  double* out; long long* clk; cudaMalloc(&out, 8192); cudaMalloc(&clk, 8);
  const char* names[] = {"DFMA", "rsqrt(double)", "sqrt(double)", "div(double)", "__syncthreads(512)+DADD", "rsqrtf+cvt"};
  const int iters = 2048;
  long long h;
#define RUN(OP, T) chain<OP><<<1, T>>>(out, clk, 1.25, iters); chain<OP><<<1, T>>>(out, clk, 1.25, iters); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("%-26s %6.1f clk/op (%d threads)\n", names[OP], (double)h / iters, T);
  RUN(0, 32) RUN(1, 32) RUN(2, 32) RUN(3, 32) RUN(4, 512) RUN(5, 32) RUN(0, 512) RUN(1, 512)
  return 0;
}
// --- barrier behaviour under role divergence (the shape of a Jacobi round) ---
// OP 0: every warp, barrier only; 1: warp 0 runs a 32-DFMA dependent chain (~290 clocks), the others go straight to the
// barrier; 2: as 1, plus every warp reads and writes shared memory after the barrier and a second barrier follows.
template <int OP>
__global__ void rounds(double* out, long long* clk, int iters) {
  __shared__ double sh[1024];
  double x = 1.0 + threadIdx.x * 1e-12;
  sh[threadIdx.x & 1023] = x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (OP >= 1 && threadIdx.x < 32) {
#pragma unroll
      for (int k = 0; k < 32; k++) x = fma(x, 0.999999, 1e-7);
      sh[threadIdx.x] = x;
    }
    __syncthreads();
    if (OP == 2) {
      const double y = sh[(threadIdx.x * 7) & 31];
      sh[32 + (threadIdx.x & 511)] = fma(y, 0.5, x);
      __syncthreads();
    }
  }
  long long t1 = clock64();
  out[threadIdx.x] = x + sh[threadIdx.x & 1023];
  if (threadIdx.x == 0) clk[0] = t1 - t0;
}
struct RunRounds {
  RunRounds() {
    double* out; long long* clk; cudaMalloc(&out, 8192); cudaMalloc(&clk, 8);
    long long h; const int iters = 2000;
    rounds<0><<<1, 512>>>(out, clk, iters); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("round: barrier only                         %6.1f clk\n", (double)h / iters);
    rounds<1><<<1, 512>>>(out, clk, iters); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("round: warp 0 chain of 32 DFMA + barrier       %6.1f clk\n", (double)h / iters);
    rounds<2><<<1, 512>>>(out, clk, iters); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("round: + shared load/store + second barrier   %6.1f clk\n", (double)h / iters);
    cudaDeviceSynchronize();
  }
} run_rounds_at_exit_of_static_init;

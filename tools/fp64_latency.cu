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

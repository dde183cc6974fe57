// Development probe: cost of a cluster barrier (8 CTAs x 544 threads) with and without remote shared-memory traffic.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
template <int OP>
__global__ void __cluster_dims__(8, 1, 1) probe(double* out, long long* clk, int iters) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ double sm[];
  const unsigned rank = cluster.block_rank();
  double* remote = cluster.map_shared_rank(sm, (rank + 1) & 7);
  sm[threadIdx.x] = threadIdx.x;
  cluster.sync();
  double x = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    if (OP >= 1) remote[threadIdx.x] = x + i;          // remote store
    if (OP >= 2) x += remote[(threadIdx.x + 1) & 511]; // remote load
    cluster.sync();
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + sm[threadIdx.x];
  if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}
int main() {
  double* out; long long* clk; cudaMalloc(&out, 8 * 544 * 8 * 2); cudaMalloc(&clk, 8);
  long long h; const int iters = 2000; const size_t smem = 150 * 1024;
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe<0><<<16, 544, smem>>>(out, clk, iters); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("cluster.sync only:            %7.1f clk  (%s)\n", (double)h / iters, cudaGetErrorString(cudaGetLastError()));
  probe<1><<<16, 544, smem>>>(out, clk, iters); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("+ remote store per thread:    %7.1f clk\n", (double)h / iters);
  probe<2><<<16, 544, smem>>>(out, clk, iters); cudaMemcpy(&h, clk, 8, cudaMemcpyDeviceToHost); printf("+ remote load per thread:     %7.1f clk\n", (double)h / iters);
  return 0;
}

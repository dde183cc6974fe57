// Row f1: the two reductions of the CDIIS step that the host framework runs as degenerate GEMMs
// (grad_dft/evaluate.py:1041-1205).  Inside the graph-replayed SCF iteration of a small molecule the iteration time IS
// the sum of its kernels' durations, and these two were 21 us each at n = 43 (a 10 x 10 output with K = n^2 on ONE CTA
// of a 32 x 64-tile GEMM; a 1 x n^2 output with K = 10):
//   gram[s][i][j] = sum_kl e[i][s][k][l] e[j][s][k][l]          "iskl,jskl->sij"   (evaluate.py:1165)
//   out[s][k][l]  = sum_i x[s][i] f[i][s][k][l]                 "si,isjk->sjk"     (evaluate.py:1198)
// One warp per gram entry (fixed lane-strided order + the warp tree: deterministic), one thread per output element.
#include "common.cuh"

namespace gdft {

// BORDERED: write the CDIIS matrix B[2, m+1, m+1] of evaluate.py:1167-1181 instead of the bare Gram matrix:
// B[0,0] = 0, B[0,1+i] = B[1+i,0] = live_i, B[1+i,1+j] = G_ij, and B[1+i,1+i] = 1 for the slots that are not live yet
// (live_i = i <= cycle).
template <bool BORDERED>
__global__ void __launch_bounds__(256) diis_gram_kernel(int m, int64_t nn, int cycle, const double* __restrict__ e, double* __restrict__ gram) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int total = 2 * m * m;
  if (warp >= total) return;
  const int s = warp / (m * m), rem = warp - s * m * m, i = rem / m, j = rem - i * m;
  if (j < i) return;  // symmetric: the (j, i) entry is written by the (i, j) warp
  const double* a = e + ((size_t)i * 2 + s) * nn;
  const double* b = e + ((size_t)j * 2 + s) * nn;
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  int64_t k = lane;
  for (; k + 96 < nn; k += 128) {
#pragma unroll
    for (int u = 0; u < 4; u++) acc[u] = fma(a[k + 32 * u], b[k + 32 * u], acc[u]);
  }
  for (; k < nn; k += 32) acc[0] = fma(a[k], b[k], acc[0]);
  const double v = warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
  if (lane == 0) {
    if (BORDERED) {
      const int mb = m + 1;
      double* B = gram + (size_t)s * mb * mb;
      const bool live = i <= cycle;
      B[(1 + i) * mb + 1 + j] = (i == j && !live) ? 1.0 : v;
      B[(1 + j) * mb + 1 + i] = (i == j && !live) ? 1.0 : v;
      if (i == j) {
        B[1 + i] = live ? 1.0 : 0.0;
        B[(1 + i) * mb] = live ? 1.0 : 0.0;
        if (i == 0) B[0] = 0.0;
      }
    } else {
      gram[((size_t)s * m + i) * m + j] = v;
      gram[((size_t)s * m + j) * m + i] = v;
    }
  }
}

__global__ void __launch_bounds__(256) diis_combine_kernel(int m, int64_t nn, const double* __restrict__ x, const double* __restrict__ f,
                                                          double* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 2 * nn) return;
  const int s = (int)(idx / nn);
  const int64_t kl = idx - (int64_t)s * nn;
  double acc = 0.0;
  for (int i = 0; i < m; i++) acc = fma(x[s * m + i], f[((size_t)i * 2 + s) * nn + kl], acc);
  out[idx] = acc;
}

}  // namespace gdft

using namespace gdft;

extern "C" int gdft_diis_gram(gdft_stream_t stream, int m, int64_t n, const double* err /*[m,2,n,n]*/, double* gram /*[2,m,m]*/) {
  if (m <= 0 || m > 64 || n <= 0 || n > 32768) return GDFT_BAD_SHAPE;
  if (!err || !gram) return GDFT_BAD_ARGUMENT;
  const int warps = 2 * m * m;
  diis_gram_kernel<false><<<(warps * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(m, n * n, 0, err, gram);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_diis_matrix(gdft_stream_t stream, int m, int64_t n, int cycle, const double* err /*[m,2,n,n]*/,
                                double* B /*[2,m+1,m+1]*/) {
  if (m <= 0 || m > 64 || n <= 0 || n > 32768) return GDFT_BAD_SHAPE;
  if (!err || !B) return GDFT_BAD_ARGUMENT;
  const int warps = 2 * m * m;
  diis_gram_kernel<true><<<(warps * 32 + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(m, n * n, cycle, err, B);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

extern "C" int gdft_diis_combine(gdft_stream_t stream, int m, int64_t n, const double* x /*[2,m]*/, const double* fock_vec /*[m,2,n,n]*/,
                                 double* out /*[2,n,n]*/) {
  if (m <= 0 || m > 64 || n <= 0 || n > 32768) return GDFT_BAD_SHAPE;
  if (!x || !fock_vec || !out) return GDFT_BAD_ARGUMENT;
  const int64_t total = 2 * n * n;
  diis_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(m, n * n, x, fock_vec, out);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}

// abs_clip (grad_dft/molecule.py:687-689): out = |src| > thr ? x : 0.  With x = src it is the clip itself; with x = the
// incoming cotangent it is its VJP (and, applied again, the VJP of that).  The host-framework composite is four launches
// (abs, compare, zeros_like, where) per call and three more in its reverse pass; the predictor calls it on the
// densities and on the Fock matrix of every build.
namespace gdft {
__global__ void __launch_bounds__(256) abs_clip_kernel(int64_t count, const double* __restrict__ x, const double* __restrict__ src, double thr,
                                                      double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = fabs(src[i]) > thr ? x[i] : 0.0;
}
}  // namespace gdft

extern "C" int gdft_abs_clip(gdft_stream_t stream, int64_t count, const double* x, const double* src, double thr, double* out) {
  if (count < 0) return GDFT_BAD_SHAPE;
  if (count == 0) return GDFT_OK;
  if (!x || !src || !out) return GDFT_BAD_ARGUMENT;
  const unsigned grid = (unsigned)imin64((count + 255) / 256, 148 * 16);
  gdft::abs_clip_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(count, x, src, thr, out);
  GDFT_LAUNCH_CHECK();
  return GDFT_OK;
}
